"""CPU tests of the drop-in boundary: libmpm_b200.so loads, exports every symbol include/mpm_b200.h
declares, its structs have the layout the reference's blittable structs have, and its presets carry the
reference's constants.  No compute is called without a GPU; and without one the library must FAIL LOUDLY."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mpm_b200.h")).read()
    return sorted(set(re.findall(r"MPM_API\s+[\w\s\*]+?\b(mpm_\w+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    import mpm_b200
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libmpm_b200.so does not export {n}"
    assert sorted(mpm_b200.EXPORTS) == names, "python binding and header disagree on the export list"


def test_abi_version_and_struct_layout(lib):
    import mpm_b200
    assert lib.mpm_abi_version() == 1
    assert C.sizeof(mpm_b200.MpmParams) == 4 * 35  # all members 4 bytes, no padding
    assert mpm_b200.default_params("3d_gpu").struct_size == C.sizeof(mpm_b200.MpmParams)  # C side agrees
    assert mpm_b200.PARTICLE80.itemsize == 80     # MLSMPM3DFluidMultithreadGPU.cs:8-22 (std430, 80 B)
    offs = {k: mpm_b200.PARTICLE80.fields[k][1] for k in ("pos", "vel", "mass", "C_x", "C_y", "C_z")}
    assert offs == {"pos": 0, "vel": 16, "mass": 28, "C_x": 32, "C_y": 48, "C_z": 64}


@pytest.mark.parametrize("name", helpers.VARIANT_NAMES)
def test_presets_carry_the_reference_constants(lib, name):
    import mpm_b200
    p = mpm_b200.default_params(name)
    o = orc.variant(name)
    assert p.dim == o.dim and list(p.grid_size) == list(o.grid)
    for f in helpers._SHARED:
        assert getattr(p, f) == getattr(o, f), f
    assert list(p.sphere_pos) == list(o.sphere_pos)


def test_bad_params_are_rejected_before_touching_cuda(lib):
    import mpm_b200
    p = mpm_b200.default_params("3d_gpu")
    p.struct_size = 12
    h = C.c_void_p()
    assert lib.mpm_create(C.byref(p), 1000, 0, C.byref(h)) == mpm_b200.ERR_INVALID
    assert b"struct_size" in lib.mpm_last_error(None)
    p = mpm_b200.default_params("3d_gpu"); p.dim = 4
    assert lib.mpm_create(C.byref(p), 1000, 0, C.byref(h)) == mpm_b200.ERR_INVALID
    p = mpm_b200.default_params("3d_fixed"); p.bc_mode = 1
    assert lib.mpm_create(C.byref(p), 1000, 0, C.byref(h)) == mpm_b200.ERR_INVALID


def test_no_gpu_means_loud_failure_not_cpu_fallback(lib):
    import mpm_b200
    if lib.mpm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = mpm_b200.default_params("3d_gpu")
    with pytest.raises(mpm_b200.MpmError) as e:
        mpm_b200.Solver(p, 1000)
    assert e.value.code == mpm_b200.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference algorithm on the host cores) needs no GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["metric"] == "particle-steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_committed_traffic_table_has_the_kernels_the_bench_reports():
    """`bench.py` fills `roofline.traffic` / `kernels[*].traffic` from profiles/r2/traffic_c4.json (DRAM bytes per launch
    from one committed `ncu --set full` capture): the table must name the three stencil kernels, cite its source file, that
    file must be in the repo, and the cell kernels must not have changed since the commit the capture was taken at (where a
    git history is available: the GPU box runs from a snapshot without one)."""
    import json
    import subprocess
    with open(os.path.join(ROOT, "profiles", "r2", "traffic_c4.json")) as f:
        t = json.load(f)
    for k in ("k_p2g1_cell", "k_p2g2_cell", "k_g2p_cell"):
        assert t["kernels"][k]["dram_bytes"] > 1e9 and t["kernels"][k]["ncu_duration_s"] > 0
        assert 1.0 <= t["kernels"][k]["dram_over_algorithmic"] < 1.5
    src = t["source"].split(" ")[0]
    assert os.path.exists(os.path.join(ROOT, src)), src
    if os.path.isdir(os.path.join(ROOT, ".git")):
        have = subprocess.run(["git", "-C", ROOT, "cat-file", "-e", t["commit"] + "^{commit}"], capture_output=True)
        if have.returncode == 0:
            changed = subprocess.run(["git", "-C", ROOT, "diff", "--quiet", t["commit"], "--", "mls-mpm-godot_b200/csrc/mpm_kernels_cell.cu"])
            assert changed.returncode == 0, "the cell kernels changed after the committed ncu capture: re-capture profiles/r2/traffic_c4.json"


def test_every_kernel_launched_with_the_pdl_attribute_waits_first():
    """A kernel launched through launch_pdl() (programmatic stream serialization) may be scheduled while the kernel before it
    is still running: it must execute griddepcontrol.wait -- pdl_prologue() or pdl_wait() -- before anything else.  Static check
    over the sources: every kernel name that appears in a launch_pdl(...) call has one of the two as the first statement
    of its body."""
    import glob
    import re
    src = {f: open(f).read() for f in glob.glob(os.path.join(ROOT, "mls-mpm-godot_b200", "csrc", "*.cu*"))}
    text = "\n".join(src.values())
    launched = set(re.findall(r"launch_pdl(?:<[^>]*>)?\(\s*([A-Za-z_0-9]+)", text))
    launched |= set(re.findall(r"LAUNCH_CELL\(\s*([A-Za-z_0-9]+)", text))  # (the macro launches its first argument)
    launched -= {"KERNEL", "void"}      # the macro's own parameter; launch_pdl's definition
    aliases = dict(re.findall(r"constexpr auto ([A-Za-z_0-9]+) = ([A-Za-z_0-9]+)<", text))  # k_g2p_cell_comm = k_g2p_cell<...>
    kernels = {aliases.get(k, k) for k in launched}
    assert len(kernels) >= 15, kernels
    for k in sorted(kernels):
        m = None
        for m in re.finditer(r"__global__[^;{]*?\b" + k + r"\s*\(", text):
            i, depth = m.end(), 1
            while depth:
                depth += {"(": 1, ")": -1}.get(text[i], 0)
                i += 1
            j = text.index("{", i)
            if text[i:j].strip():
                continue  # a declaration
            body = text[j + 1:j + 200].lstrip()
            assert body.startswith("pdl_prologue();") or body.startswith("pdl_wait();"), f"{k} does not wait first"
            break
        assert m is not None, f"kernel {k} not found"


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "example")
    lib_dir = os.path.join(ROOT, "mls-mpm-godot_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(lib_dir, "host", "example.c"), "-L" + lib_dir, "-lmpm_b200", "-Wl,-rpath," + lib_dir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/mpm_b200.h is the drop-in boundary: a C99 program (host/example.c, the shape of the reference's _Ready /
    _Process) compiles against it without a warning, links to libmpm_b200.so and -- without a GPU -- reports the ABI version
    and the shipping scene's parameters and stops, because there is no CPU path."""
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi 1" in r.stdout and "grid 64 x 64 x 64" in r.stdout
    import mpm_b200
    if mpm_b200.load().mpm_device_count() == 0:
        assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_host_runs_the_shipping_scene(tmp_path):
    """The same C program on a GPU: InitialiseSim of the shipping scene (54^3 particles), three frames of set_sphere ->
    step(2) -> positions."""
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "frame 2: 157464 particles" in r.stdout, r.stdout
