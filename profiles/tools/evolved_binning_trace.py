"""C4 dam-break: per step from step 90 on, the binning's verdict (far movers, stable or atomic order) and the phase times."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + '/mls-mpm-godot_b200', ROOT + '/tests']
import mpm_b200
p = mpm_b200.default_params("3d_gpu", grid=(256, 256, 256), interaction=0, kernel_path=3, math_mode=1)
with mpm_b200.Solver(p, 32768000) as s:
    s.initialise_sim((4, 4, 4), (164, 164, 164), 0.5)
    for start in (0, 20, 40, 60, 80, 100, 120, 140):
        s.step(start - s.stats().steps); s.sync()
        s.set_timing(1)
        prev = s.stats().unordered_binnings
        for k in range(6):
            s.step(1); s.sync()
            st = s.stats()
            print(f"step {st.steps:4d} far={st.far_movers:8d} unordered={'yes' if st.unordered_binnings > prev else 'no '} sort={st.ms_sort:.3f} "
                  f"p2g1={st.ms_p2g1:.3f} p2g2={st.ms_p2g2:.3f} g2p={st.ms_g2p:.3f} clear+update={st.ms_clear + st.ms_update:.3f} step={st.ms_step:.3f} cells={st.num_cells}", flush=True)
            prev = st.unordered_binnings
        s.set_timing(0)
