// MpmB200.cs -- P/Invoke binding of libmpm_b200.so (include/mpm_b200.h) plus a Godot node that exposes the
// surface of the reference's GPU solver node (mls-mpm/3d/fluid_multithread_gpu/MLSMPM3DFluidMultithreadGPU.cs)
// on top of it.  Drop this file next to the reference's solver scripts, put libmpm_b200.so on the library
// path, and attach MLSMPM3DFluidB200 where MLSMPM3DFluidMultithreadGPU was attached.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no dotnet/godot.  The same ABI is exercised from Python
// ctypes (mls-mpm-godot_b200/mpm_b200/__init__.py) and tests/test_abi.py checks that the struct layouts below
// (sizes, offsets) are the ones the library uses.  Citations "H:n" are lines of the reference GPU node.
using Godot;
using System;
using System.Runtime.InteropServices;

namespace MpmB200
{
    // ---- blittable records: identical to the reference's (H:8-33), so Particle[] / Cell[] pin and pass directly
    [StructLayout(LayoutKind.Sequential)]
    public struct Particle            // 80 bytes, std430 (H:8-22)
    {
        public Vector3 pos; public float padding_pos;
        public Vector3 vel; public float mass;
        public Vector3 C_x; public float padding_c_x;   // column 0 of C (Basis.X)
        public Vector3 C_y; public float padding_c_y;
        public Vector3 C_z; public float padding_c_z;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct Cell                // 16 bytes (H:25-32): int32 x fixed_point_mult
    {
        public int vel_x, vel_y, vel_z, mass;
    }

    // union of the reference's push-constant blocks (H:444-503) + the constants its five solver copies differ in
    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct MpmParams    // 35 x 4 bytes, no padding (MpmParams in mpm_b200.h)
    {
        public int struct_size, dim;
        public fixed int grid_size[3];
        public float dt, gravity, rest_density, dynamic_viscosity, eos_stiffness, eos_power;
        public int grid_mode, fixed_point_mult, stress_form, eq16_order, bc_mode, bc_hi_off;
        public float bc_friction, clamp_min, clamp_max_off, wall_min, wall_max_off, wall_gain;
        public int interaction;
        public fixed float sphere_pos[3];
        public float sphere_radius;
        public fixed float mouse_pos[2];
        public float mouse_radius;
        public int math_mode, kernel_path, sort_interval, overflow_check;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MpmStats
    {
        public long num_particles, num_cells, steps, kernel_launches;
        public float ms_sort, ms_clear, ms_p2g1, ms_p2g2, ms_update, ms_g2p, ms_exchange, ms_step;
        public int kernel_path, overflow, rank, world;
        public long local_particles, migrated, slab_jump_clamps, unordered_binnings, far_movers, halo_peer_exchanges;
        public float ms_halo_mass, ms_halo_momentum, ms_migration;
        public int reserved0;
    }

    public static unsafe class Native
    {
        const string Lib = "mpm_b200";
        const CallingConvention CC = CallingConvention.Cdecl;
        public const int VARIANT_3D_GPU = 4, MATH_STRICT = 0, MATH_FAST = 1;
        public const int PATH_AUTO = 0, PATH_REFERENCE = 1, PATH_TILED = 2, PATH_CELL = 3;
        public const int ERR_DOMAIN = 6;

        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_abi_version();
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_device_count();
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_default_params(int variant, MpmParams* p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_create(MpmParams* p, long max_particles, int device, out IntPtr solver);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_destroy(IntPtr s);
        [DllImport(Lib, CallingConvention = CC)] public static extern IntPtr mpm_last_error(IntPtr s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_set_params(IntPtr s, MpmParams* p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_params(IntPtr s, MpmParams* p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_set_sphere(IntPtr s, float* pos3);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_set_colliders(IntPtr s, float* xyzr, int count);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_init_block(IntPtr s, float* lo3, float* hi3, float spacing);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_add_block(IntPtr s, float* lo3, float* hi3, float spacing);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_upload_particles(IntPtr s, [In] Particle[] ps, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_download_particles(IntPtr s, [Out] Particle[] ps, long cap);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_download_grid(IntPtr s, [Out] Cell[] cells, long cap);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_save_state(IntPtr s, [MarshalAs(UnmanagedType.LPUTF8Str)] string path);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_load_state(IntPtr s, [MarshalAs(UnmanagedType.LPUTF8Str)] string path);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_step(IntPtr s, int iterations);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_sync(IntPtr s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_run_phase(IntPtr s, int phase);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_positions(IntPtr s, IntPtr dst4, long cap, out IntPtr device_ptr, out uint tex_width);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_export_positions(IntPtr s, out int fd, out ulong bytes, out uint tex_width);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_positions_async(IntPtr s, IntPtr dst4, long cap);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_positions_q16_async(IntPtr s, IntPtr dst4, long cap);  // 4 x uint16 per particle
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_wait_positions(IntPtr s);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_num_particles(IntPtr s, out long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_set_timing(IntPtr s, int enabled);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_stats(IntPtr s, out MpmStats st);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_host_alloc(long bytes, out IntPtr p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_host_free(IntPtr p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_get_stream(IntPtr s, out IntPtr stream);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_upload_particles_soa(IntPtr s, float* pos, float* vel, float* C, float* mass, long n);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_download_particles_soa(IntPtr s, float* pos, float* vel, float* C, float* mass, long cap);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_debug_last_sort(IntPtr s, uint* keys_before, uint* perm, long cap);
        // multi-GPU x-slabs (no reference counterpart): NCCL (one process per GPU) or LOCAL (k solvers in this process)
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_unique_id(byte* id128);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_init(IntPtr s, byte* id128, int rank, int world);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_local_hub_create(int world, out IntPtr hub);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_local_hub_destroy(IntPtr hub);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_init_local(IntPtr s, IntPtr hub, int rank, int world);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_rebalance(IntPtr s, int max_shift);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_rebalance_weighted(IntPtr s, int max_shift, float cost_per_particle);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_comm_slab(IntPtr s, out int x0, out int x1, out int gx0, out int nxl);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_download_ids(IntPtr s, uint* ids, long cap);
        [DllImport(Lib, CallingConvention = CC)] public static extern int mpm_slab_cuts(long* hist, int rx, int world, int min_width, int* cuts);

        public static void Check(int rc, IntPtr s)
        {
            if (rc != 0) throw new InvalidOperationException($"mpm_b200 error {rc}: {Marshal.PtrToStringUTF8(mpm_last_error(s))}");
        }
    }

    /// Same exported parameters, lifecycle and outputs as MLSMPM3DFluidMultithreadGPU (H:54-84, 158-251, 546-616).
    public unsafe partial class MLSMPM3DFluidB200 : Node3D
    {
        Vector3I grid_size = new Vector3I(64, 64, 64);          // H:43
        const int max_particle_count = 300000;                  // H:46
        IntPtr solver = IntPtr.Zero;
        MpmParams prm;
        IntPtr host_positions = IntPtr.Zero;                    // pinned float4[num_particles]
        long num_particles;
        public uint particle_pos_tex_width;                      // H:196
        ImageTexture particle_pos_tex;                          // replaces the rgba32f RD texture (H:342-355)
        Image particle_pos_img;
        MultiMeshInstance3D multi_mesh_instance;

        [Export(PropertyHint.Range, "0.0f,0.4f,")]
        float Dt { get => prm.dt; set { prm.dt = Math.Clamp(value, 0.0f, 0.4f); UpdatePushConstants(); } }   // H:57-67
        [Export] int sim_iterations = 2;                                                                   // H:69
        [Export] public float gravity { get => prm.gravity; set { prm.gravity = value; } }                 // H:71; the UI calls UpdatePushConstants() after setting it
        [Export] float rest_density { get => prm.rest_density; set { prm.rest_density = value; } }         // H:76
        [Export] float dynamic_viscosity { get => prm.dynamic_viscosity; set { prm.dynamic_viscosity = value; } }
        [Export] float eos_stiffness { get => prm.eos_stiffness; set { prm.eos_stiffness = value; } }      // H:82
        [Export] float eos_power { get => prm.eos_power; set { prm.eos_power = value; } }                  // H:84
        [Export] PhysicsBody3D sphere_body;                                                                // H:90
        // B200 solver controls (no reference counterpart).  The node defaults to the fast path -- MATH_FAST + PATH_AUTO
        // resolves to the cell kernels (11.5 G particle-steps/s on the 32.8 M-particle benchmark scene, results within the
        // FAST tolerance of the reference algorithm and reproducible run to run); MATH_STRICT gives the bit-exact tiled
        // kernels (5.3 G).  Both are fixed when the solver is created (_Ready).
        [Export(PropertyHint.Enum, "Strict (bit-exact),Fast")] int math_mode { get => prm.math_mode; set { prm.math_mode = value; } }
        [Export(PropertyHint.Enum, "Auto,Reference-shaped,Tiled,Cell")] int kernel_path { get => prm.kernel_path; set { prm.kernel_path = value; } }
        /// File descriptor of the device allocation that holds the (x, y, z, |v|) array (-1 if the driver cannot export):
        /// a GDExtension renderer imports it once through VK_KHR_external_memory_fd and samples it with no host round trip.
        public int particle_pos_fd = -1;
        public ulong particle_pos_bytes;
        /// false: skip the host copy in _Process (a renderer reads the exported allocation instead)
        [Export] bool copy_positions_to_host = true;

        public MLSMPM3DFluidB200()
        {
            fixed (MpmParams* p = &prm) Native.mpm_default_params(Native.VARIANT_3D_GPU, p);
            prm.math_mode = Native.MATH_FAST;      // (mpm_default_params gives MATH_STRICT, the conservative library default)
            prm.kernel_path = Native.PATH_AUTO;
        }

        public override void _Ready()                                            // H:158-207
        {
            multi_mesh_instance = GetNode<MultiMeshInstance3D>("MultiMeshInstance3D");
            prm.grid_size[0] = grid_size.X; prm.grid_size[1] = grid_size.Y; prm.grid_size[2] = grid_size.Z;
            if (sphere_body != null) SetSphere(sphere_body.GlobalPosition); else GD.PrintErr("sphere_body not set");
            fixed (MpmParams* p = &prm) Native.Check(Native.mpm_create(p, max_particle_count, 0, out solver), IntPtr.Zero);
            InitialiseSim();
            if (Native.mpm_export_positions(solver, out particle_pos_fd, out particle_pos_bytes, out _) != 0) particle_pos_fd = -1;
            particle_pos_tex_width = (uint)Mathf.Sqrt(num_particles) + 1;
            Native.Check(Native.mpm_host_alloc(16L * particle_pos_tex_width * particle_pos_tex_width, out host_positions), solver);
            particle_pos_img = Image.CreateEmpty((int)particle_pos_tex_width, (int)particle_pos_tex_width, false, Image.Format.Rgbaf);
            particle_pos_tex = ImageTexture.CreateFromImage(particle_pos_img);
            if (multi_mesh_instance != null)
            {
                multi_mesh_instance.Multimesh.InstanceCount = (int)num_particles;
                (multi_mesh_instance.MaterialOverride as ShaderMaterial)?.SetShaderParameter("particle_pos_tex", particle_pos_tex);  // H:404-412
            }
            var global_node = GetTree().Root.GetNode<Node>("Global");             // H:203-207
            global_node.Set("particle_count", num_particles);
            global_node.Set("particle_pos_texture", particle_pos_tex);
            global_node.Set("particle_pos_texture_width", particle_pos_tex_width);
            global_node.Set("current_simulator", this);
        }

        void InitialiseSim()                                                      // H:654-707: centred 32^3 box, spacing 0.6
        {
            const float spacing = 0.6f; const int box = 32;
            float* lo = stackalloc float[3]; float* hi = stackalloc float[3];
            lo[0] = grid_size.X / 2 - box / 2; lo[1] = grid_size.Y / 2 - box / 2; lo[2] = grid_size.Z / 2 - box / 2;
            hi[0] = lo[0] + box; hi[1] = lo[1] + box; hi[2] = lo[2] + box;
            Native.Check(Native.mpm_init_block(solver, lo, hi, spacing), solver);
            Native.Check(Native.mpm_num_particles(solver, out num_particles), solver);
            GD.Print("num_particles: ", num_particles);
        }

        public void UpdatePushConstants()                                         // H:444-503; also called by main_ui.tscn:70-72
        {
            if (solver == IntPtr.Zero) return;
            fixed (MpmParams* p = &prm) Native.Check(Native.mpm_set_params(solver, p), solver);
        }

        void SetSphere(Vector3 pos)                                               // HandleMouseInteraction, H:618-642
        {
            prm.sphere_pos[0] = pos.X; prm.sphere_pos[1] = pos.Y; prm.sphere_pos[2] = pos.Z;
            if (solver == IntPtr.Zero) return;
            float* a = stackalloc float[3]; a[0] = pos.X; a[1] = pos.Y; a[2] = pos.Z;
            Native.Check(Native.mpm_set_sphere(solver, a), solver);
        }

        public override void _Process(double delta)                               // H:234-251
        {
            if (sphere_body != null) SetSphere(sphere_body.GlobalPosition);
            Native.Check(Native.mpm_step(solver, sim_iterations), solver);        // sim_iterations x (clear, P2G_1, P2G_2, update, G2P)
            // particle_pos_tex hand-off (g2p.glsl:149-150): (x, y, z, |v|) at texel (i % W, i / W).
            // Zero-copy: a renderer that imported particle_pos_fd only needs the device array refreshed (dst = null).
            if (!copy_positions_to_host)
            {
                Native.Check(Native.mpm_get_positions(solver, IntPtr.Zero, 0, out _, out _), solver);
                Native.Check(Native.mpm_sync(solver), solver);
                return;
            }
            // Portable path: through pinned host memory into an ImageTexture.
            Native.Check(Native.mpm_get_positions(solver, host_positions, num_particles, out _, out _), solver);
            var bytes = new byte[16 * particle_pos_tex_width * particle_pos_tex_width];
            Marshal.Copy(host_positions, bytes, 0, (int)(16 * num_particles));
            particle_pos_img.SetData((int)particle_pos_tex_width, (int)particle_pos_tex_width, false, Image.Format.Rgbaf, bytes);
            particle_pos_tex.Update(particle_pos_img);
        }

        public void TurnOnVisualisation() { multi_mesh_instance.Visible = true; }  // H:253-261
        public void TurnOffVisualisation() { multi_mesh_instance.Visible = false; }

        public override void _Notification(int what)                              // CleanupGpu, H:546-616, 709-715
        {
            if (what == NotificationPredelete)
            {
                if (host_positions != IntPtr.Zero) { Native.mpm_host_free(host_positions); host_positions = IntPtr.Zero; }
                if (solver != IntPtr.Zero) { Native.mpm_destroy(solver); solver = IntPtr.Zero; }
            }
        }
    }
}
