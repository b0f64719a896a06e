/*
 * mpm_oracle.c -- TEST INFRASTRUCTURE ONLY (see mpm_oracle.h for the contract and the citation keys
 * F / X / D / M / H).  PARITY UNPINNED: the reference has no golden vectors and cannot run here.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -pthread -fPIC -shared (oracle/Makefile).
 * Every floating-point expression below is written in the source order of the C# statement it restates,
 * one IEEE binary32 operation per C operator; -ffp-contract=off forbids FMA fusion (RyuJIT never fuses).
 */
#include "mpm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---------------------------------------------------------------- tiny Parallel.For (F:229) on pthreads */

typedef void (*range_fn)(void* ctx, int64_t lo, int64_t hi);
typedef struct { range_fn fn; void* ctx; int64_t lo, hi; } PfTask;
static void* pf_thread(void* a) { PfTask* t = (PfTask*)a; t->fn(t->ctx, t->lo, t->hi); return NULL; }

static void parallel_for(int nt, int64_t n, range_fn fn, void* ctx)
{
    if (nt <= 1 || n < 2048) { fn(ctx, 0, n); return; }
    if (nt > 256) nt = 256;
    pthread_t th[256];
    PfTask task[256];
    int64_t chunk = (n + nt - 1) / nt;
    int started = 0;
    for (int t = 0; t < nt; ++t) {
        int64_t lo = t * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo >= hi) break;
        task[t].fn = fn; task[t].ctx = ctx; task[t].lo = lo; task[t].hi = hi;
        if (pthread_create(&th[t], NULL, pf_thread, &task[t]) != 0) { fn(ctx, lo, hi); th[t] = 0; }
        started = t + 1;
    }
    for (int t = 0; t < started; ++t) if (th[t]) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------- scalar helpers */

/* X:151-154  (int)(floating_point * fixed_point_mult): int mult widens to float, fp32 product, truncate */
int32_t orc_encode_fixed(float f, int32_t mult) { return (int32_t)(f * (float)mult); }
/* X:156-159  (float)(fixed_point) / fixed_point_mult */
float orc_decode_fixed(int32_t i, int32_t mult) { return (float)i / (float)mult; }

float orc_pow(int32_t mode, float x, float y)
{
    if (mode == ORC_POW_LIBM_POWF) return powf(x, y); /* F:331 Mathf.Pow -> MathF.Pow -> CRT powf */
    float yi = truncf(y);
    if (yi == y && y >= 1.0f && y <= 64.0f) {
        double xd = (double)x, r = xd;
        int n = (int)y;
        for (int k = 1; k < n; ++k) r = r * xd;
        return (float)r;
    }
    return (float)pow((double)x, (double)y);
}

/* F:259-263 quadratic B-spline weights for one axis */
static inline int axis_weights(float p, float w[3])
{
    int c = (int)p;                        /* (Vector3I)p.pos : truncation */
    float cd = (p - (float)c) - 0.5f;      /* (p.pos - cell_idx) - 0.5 */
    w[0] = 0.5f * ((0.5f - cd) * (0.5f - cd));
    w[1] = 0.75f - (cd * cd);
    w[2] = 0.5f * ((0.5f + cd) * (0.5f + cd));
    return c;
}

static inline int64_t num_cells(const OrcParams* P)
{
    return (int64_t)P->grid[0] * P->grid[1] * (P->dim == 3 ? P->grid[2] : 1);
}

/* ---------------------------------------------------------------- scene */

/* F:136-146 / H:661-671: for (float i = lo; i < hi; i += spacing) nested x, y, z */
int64_t orc_init_block(int32_t dim, const float lo[3], const float hi[3], float spacing, float* pos,
                       int64_t cap)
{
    int64_t n = 0;
    for (float i = lo[0]; i < hi[0]; i += spacing)
        for (float j = lo[1]; j < hi[1]; j += spacing) {
            if (dim == 2) {
                if (pos && n < cap) { pos[3 * n] = i; pos[3 * n + 1] = j; pos[3 * n + 2] = 0.0f; }
                ++n;
                continue;
            }
            for (float k = lo[2]; k < hi[2]; k += spacing) {
                if (pos && n < cap) { pos[3 * n] = i; pos[3 * n + 1] = j; pos[3 * n + 2] = k; }
                ++n;
            }
        }
    return n;
}

/* ---------------------------------------------------------------- ClearGrid  F:222-250, X:245-275 */

static void clear_range(void* ctx, int64_t lo, int64_t hi)
{
    int32_t* g = (int32_t*)ctx; /* +0.0f and int 0 share the all-zero bit pattern */
    for (int64_t i = lo; i < hi; ++i) { g[4 * i] = 0; g[4 * i + 1] = 0; g[4 * i + 2] = 0; g[4 * i + 3] = 0; }
}
static void clear_grid_impl(const OrcParams* P, void* grid, int nt)
{
    parallel_for(nt, num_cells(P), clear_range, grid);
}
void orc_clear_grid(const OrcParams* P, void* grid) { clear_grid_impl(P, grid, 1); }

/* ---------------------------------------------------------------- P2G_1  F:252-294, X:290-343, D:197-235 */

static inline void add_cell(const OrcParams* P, void* grid, int64_t ci, float m, float vx, float vy,
                            float vz, int with_mass, int atomic)
{
    if (P->grid_mode == ORC_GRID_FLOAT) {
        float* c = (float*)grid + 4 * ci;
        if (with_mass) c[3] += m; /* cell.mass += mass_contrib  F:286 */
        c[0] += vx;               /* cell.vel += ...            F:287 */
        c[1] += vy;
        c[2] += vz;
    } else {
        int32_t* c = (int32_t*)grid + 4 * ci;
        int32_t mult = P->fixed_point_mult;
        int32_t em = with_mass ? orc_encode_fixed(m, mult) : 0;
        int32_t ex = orc_encode_fixed(vx, mult), ey = orc_encode_fixed(vy, mult),
                ez = orc_encode_fixed(vz, mult);
        if (atomic) { /* Interlocked.Add  X:336-339 */
            if (with_mass) __atomic_fetch_add(&c[3], em, __ATOMIC_RELAXED);
            __atomic_fetch_add(&c[0], ex, __ATOMIC_RELAXED);
            __atomic_fetch_add(&c[1], ey, __ATOMIC_RELAXED);
            __atomic_fetch_add(&c[2], ez, __ATOMIC_RELAXED);
        } else { /* int adds wrap like Interlocked.Add; done unsigned to stay defined in C */
            if (with_mass) c[3] = (int32_t)((uint32_t)c[3] + (uint32_t)em);
            c[0] = (int32_t)((uint32_t)c[0] + (uint32_t)ex);
            c[1] = (int32_t)((uint32_t)c[1] + (uint32_t)ey);
            c[2] = (int32_t)((uint32_t)c[2] + (uint32_t)ez);
        }
    }
}

static void p2g1_particle(const OrcParams* P, int64_t i, const float* pos, const float* vel,
                          const float* C, const float* mass, void* grid, int atomic)
{
    const int Ry = P->grid[1], Rz = (P->dim == 3) ? P->grid[2] : 1;
    const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
    const float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
    const float* c = C + 9 * i; /* c[3*col+row] */
    const float m = mass[i];
    float wx[3], wy[3], wz[3];
    const int cx = axis_weights(px, wx), cy = axis_weights(py, wy);
    if (P->dim == 2) {
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy) {
                float weight = wx[gx] * wy[gy];                 /* D:216 */
                int nx = cx + gx - 1, ny = cy + gy - 1;         /* D:218 */
                float dx = ((float)nx - px) + 0.5f;             /* D:219 */
                float dy = ((float)ny - py) + 0.5f;
                float mass_contrib = weight * m;                /* D:222 */
                /* Transform2D * Vector2 = (X.x*v.x + Y.x*v.y, X.y*v.x + Y.y*v.y) + Origin   D:229 */
                float qx = (c[0] * dx + c[3] * dy) + 0.0f;
                float qy = (c[1] * dx + c[4] * dy) + 0.0f;
                int64_t ci = (int64_t)nx * Ry + ny;             /* D:224 */
                add_cell(P, grid, ci, mass_contrib, mass_contrib * (vx + qx), mass_contrib * (vy + qy),
                         0.0f, 1, atomic);
            }
        return;
    }
    const int cz = axis_weights(pz, wz);
    for (int gx = 0; gx < 3; ++gx)
        for (int gy = 0; gy < 3; ++gy)
            for (int gz = 0; gz < 3; ++gz) {
                float weight = wx[gx] * wy[gy] * wz[gz];        /* F:273 */
                int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1; /* F:275 */
                float dx = ((float)nx - px) + 0.5f;             /* F:276 */
                float dy = ((float)ny - py) + 0.5f;
                float dz = ((float)nz - pz) + 0.5f;
                /* Basis * Vector3 = (Row0.Dot(v), Row1.Dot(v), Row2.Dot(v))                F:277 */
                float qx = (c[0] * dx + c[3] * dy) + c[6] * dz;
                float qy = (c[1] * dx + c[4] * dy) + c[7] * dz;
                float qz = (c[2] * dx + c[5] * dy) + c[8] * dz;
                float mass_contrib = weight * m;                /* F:279 */
                int64_t ci = ((int64_t)nx * Ry + ny) * Rz + nz; /* F:282 */
                add_cell(P, grid, ci, mass_contrib, mass_contrib * (vx + qx), mass_contrib * (vy + qy),
                         mass_contrib * (vz + qz), 1, atomic);  /* F:286-287, X:318,336-339 */
            }
}

typedef struct {
    const OrcParams* P; float* pos; float* vel; float* C; const float* mass; void* grid; int atomic;
} StepCtx;

static void p2g1_range(void* ctx, int64_t lo, int64_t hi)
{
    StepCtx* s = (StepCtx*)ctx;
    for (int64_t i = lo; i < hi; ++i) p2g1_particle(s->P, i, s->pos, s->vel, s->C, s->mass, s->grid, s->atomic);
}
static void p2g1_impl(const OrcParams* P, int64_t n, const float* pos, const float* vel, const float* C,
                      const float* mass, void* grid, int nt)
{
    StepCtx s = {P, (float*)pos, (float*)vel, (float*)C, mass, grid, 0};
    if (P->grid_mode == ORC_GRID_FIXED && nt > 1) { /* X:279-288 Parallel.For over particles */
        s.atomic = 1;
        parallel_for(nt, n, p2g1_range, &s);
    } else { /* F:254 serial */
        p2g1_range(&s, 0, n);
    }
}
void orc_p2g1(const OrcParams* P, int64_t n, const float* pos, const float* vel, const float* C,
              const float* mass, void* grid)
{
    p2g1_impl(P, n, pos, vel, C, mass, grid, 1);
}

/* ---------------------------------------------------------------- P2G_2  F:296-373, X:358-440, D:237-307 */

static inline float cell_mass(const OrcParams* P, const void* grid, int64_t ci)
{
    if (P->grid_mode == ORC_GRID_FLOAT) return ((const float*)grid)[4 * ci + 3];
    return orc_decode_fixed(((const int32_t*)grid)[4 * ci + 3], P->fixed_point_mult); /* X:383 */
}

static void p2g2_particle(const OrcParams* P, int64_t i, const float* pos, const float* C,
                          const float* mass, void* grid, int atomic)
{
    const int Ry = P->grid[1], Rz = (P->dim == 3) ? P->grid[2] : 1;
    const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
    const float* c = C + 9 * i;
    const float m = mass[i];
    const float dt = P->dt, visc = P->dynamic_viscosity;
    float wx[3], wy[3], wz[3];
    const int cx = axis_weights(px, wx), cy = axis_weights(py, wy);
    const int cz = (P->dim == 3) ? axis_weights(pz, wz) : 0;

    float density = 0.0f;                                        /* F:310 */
    if (P->dim == 2) {
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy) {
                float weight = wx[gx] * wy[gy];
                int64_t ci = (int64_t)(cx + gx - 1) * Ry + (cy + gy - 1);
                density += cell_mass(P, grid, ci) * weight;      /* D:259 */
            }
    } else {
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy)
                for (int gz = 0; gz < 3; ++gz) {
                    float weight = wx[gx] * wy[gy] * wz[gz];
                    int64_t ci = ((int64_t)(cx + gx - 1) * Ry + (cy + gy - 1)) * Rz + (cz + gz - 1);
                    density += cell_mass(P, grid, ci) * weight;  /* F:320, X:383 */
                }
    }
    float volume = m / density;                                  /* F:326 */
    /* F:331  Mathf.Max(-0.1f, eos_stiffness * (Mathf.Pow(density / rest_density, eos_power) - 1)) */
    float pw = orc_pow(P->pow_mode, density / P->rest_density, P->eos_power);
    float pr = P->eos_stiffness * (pw - 1.0f);
    float pressure = (-0.1f > pr) ? -0.1f : pr;

    if (P->dim == 2) {
        /* D:269-283: stress = diag(-p); strain = C with both off-diagonals replaced by their sum */
        float trace = c[3] + c[1];                               /* strain.Y.X + strain.X.Y */
        float sXx = c[0], sXy = trace, sYx = trace, sYy = c[4];
        float vXx = visc * sXx, vXy = visc * sXy, vYx = visc * sYx, vYy = visc * sYy; /* D:282 */
        float tXx = -pressure + vXx, tXy = 0.0f + vXy, tYx = 0.0f + vYx, tYy = -pressure + vYy; /* D:283 */
        float eXx, eXy, eYx, eYy;
        if (P->eq16_order == ORC_EQ16_DTVOL_4) {                 /* D:285  -dt * volume * stress.X * 4 */
            float s = (-dt) * volume;
            eXx = (s * tXx) * 4.0f; eXy = (s * tXy) * 4.0f; eYx = (s * tYx) * 4.0f; eYy = (s * tYy) * 4.0f;
        } else {                                                 /* M:307  -volume * 4 * stress.X * dt */
            float s = (-volume) * 4.0f;
            eXx = (s * tXx) * dt; eXy = (s * tXy) * dt; eYx = (s * tYx) * dt; eYy = (s * tYy) * dt;
        }
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy) {
                float weight = wx[gx] * wy[gy];
                int nx = cx + gx - 1, ny = cy + gy - 1;
                float dx = ((float)nx - px) + 0.5f, dy = ((float)ny - py) + 0.5f;
                /* D:300  Transform2D(e.X*w, e.Y*w, 0) * cell_dist */
                float mXx = eXx * weight, mXy = eXy * weight, mYx = eYx * weight, mYy = eYy * weight;
                float momx = (mXx * dx + mYx * dy) + 0.0f;
                float momy = (mXy * dx + mYy * dy) + 0.0f;
                int64_t ci = (int64_t)nx * Ry + ny;
                add_cell(P, grid, ci, 0.0f, momx, momy, 0.0f, 0, atomic); /* D:301 */
            }
        return;
    }

    /* F:333-345.  Columns: dudv.X = (c0,c1,c2); dudvT.X = row 0 of dudv = (c0,c3,c6) */
    float sX[3] = {c[0] + c[0], c[1] + c[3], c[2] + c[6]};
    float sY[3] = {c[3] + c[1], c[4] + c[4], c[5] + c[7]};
    float sZ[3] = {c[6] + c[2], c[7] + c[5], c[8] + c[8]};
    float tX[3] = {-pressure + sX[0] * visc, 0.0f + sX[1] * visc, 0.0f + sX[2] * visc};
    float tY[3] = {0.0f + sY[0] * visc, -pressure + sY[1] * visc, 0.0f + sY[2] * visc};
    float tZ[3] = {0.0f + sZ[0] * visc, 0.0f + sZ[1] * visc, -pressure + sZ[2] * visc};
    float eX[3], eY[3], eZ[3];
    {   /* F:347  -volume * 4 * stress.X * dt  (the only order the 3D variants use) */
        float s = (-volume) * 4.0f;
        for (int k = 0; k < 3; ++k) {
            eX[k] = (s * tX[k]) * dt; eY[k] = (s * tY[k]) * dt; eZ[k] = (s * tZ[k]) * dt;
        }
    }
    for (int gx = 0; gx < 3; ++gx)
        for (int gy = 0; gy < 3; ++gy)
            for (int gz = 0; gz < 3; ++gz) {
                float weight = wx[gx] * wy[gy] * wz[gz];         /* F:355 */
                int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1;
                float dx = ((float)nx - px) + 0.5f, dy = ((float)ny - py) + 0.5f,
                      dz = ((float)nz - pz) + 0.5f;              /* F:358 */
                /* F:364  Basis(e.X*w, e.Y*w, e.Z*w) * cell_dist, row-dot form */
                float momx = ((eX[0] * weight) * dx + (eY[0] * weight) * dy) + (eZ[0] * weight) * dz;
                float momy = ((eX[1] * weight) * dx + (eY[1] * weight) * dy) + (eZ[1] * weight) * dz;
                float momz = ((eX[2] * weight) * dx + (eY[2] * weight) * dy) + (eZ[2] * weight) * dz;
                int64_t ci = ((int64_t)nx * Ry + ny) * Rz + nz;  /* F:360 */
                add_cell(P, grid, ci, 0.0f, momx, momy, momz, 0, atomic); /* F:365, X:431-433 */
            }
}

static void p2g2_range(void* ctx, int64_t lo, int64_t hi)
{
    StepCtx* s = (StepCtx*)ctx;
    for (int64_t i = lo; i < hi; ++i) p2g2_particle(s->P, i, s->pos, s->C, s->mass, s->grid, s->atomic);
}
static void p2g2_impl(const OrcParams* P, int64_t n, const float* pos, const float* C, const float* mass,
                      void* grid, int nt)
{
    StepCtx s = {P, (float*)pos, NULL, (float*)C, mass, grid, 0};
    if (P->grid_mode == ORC_GRID_FIXED && nt > 1) { /* X:347-356 */
        s.atomic = 1;
        parallel_for(nt, n, p2g2_range, &s);
    } else { /* F:299 serial */
        p2g2_range(&s, 0, n);
    }
}
void orc_p2g2(const OrcParams* P, int64_t n, const float* pos, const float* C, const float* mass,
              void* grid)
{
    p2g2_impl(P, n, pos, C, mass, grid, 1);
}

/* ---------------------------------------------------------------- UpdateGrid  F:375-410, X:442-486, D:309-332 */

static void update_cell(const OrcParams* P, void* grid, int64_t i)
{
    const int Rx = P->grid[0], Ry = P->grid[1], Rz = (P->dim == 3) ? P->grid[2] : 1;
    int x, y, z;
    if (P->dim == 3) { x = (int)(i / Rz / Ry); y = (int)(i / Rz % Ry); z = (int)(i % Rz); } /* F:399-401 */
    else { x = (int)(i / Ry); y = (int)(i % Ry); z = 2; }                                     /* D:322-323 */
    const int hi = P->bc_hi_off;
    const int ox = (x < 2 || x > Rx - hi), oy = (y < 2 || y > Ry - hi),
              oz = (P->dim == 3) && (z < 2 || z > Rz - hi);
    if (P->grid_mode == ORC_GRID_FLOAT) {
        float* c = (float*)grid + 4 * i;
        if (c[3] > 0) {                                          /* F:392 */
            c[0] = c[0] / c[3]; c[1] = c[1] / c[3]; c[2] = c[2] / c[3]; /* F:395 */
            c[0] = c[0] + P->dt * 0.0f;                          /* F:396  dt * Vector3(0, gravity, 0) */
            c[1] = c[1] + P->dt * P->gravity;
            c[2] = c[2] + P->dt * 0.0f;
            if (P->bc_mode == ORC_BC_SLIP) {                     /* F:402-404 */
                if (ox) c[0] = 0; if (oy) c[1] = 0; if (oz) c[2] = 0;
            } else {                                             /* M:366-368 (3D: natural extension) */
                float f = P->bc_friction;
                if (ox) { c[1] = f * c[1]; c[2] = f * c[2]; c[0] = 0.0f; }
                if (oy) { c[0] = f * c[0]; c[2] = f * c[2]; c[1] = 0.0f; }
                if (oz) { c[0] = f * c[0]; c[1] = f * c[1]; c[2] = 0.0f; }
            }
        }
    } else {
        int32_t* c = (int32_t*)grid + 4 * i;
        const int32_t mult = P->fixed_point_mult;
        if (c[3] > 0) {                                          /* X:459 */
            float vx = orc_decode_fixed(c[0], mult), vy = orc_decode_fixed(c[1], mult),
                  vz = orc_decode_fixed(c[2], mult);             /* X:461-465 */
            float mm = orc_decode_fixed(c[3], mult);
            vx = vx / mm; vy = vy / mm; vz = vz / mm;            /* X:469 */
            c[0] = orc_encode_fixed(vx, mult);                   /* X:470-472 */
            c[1] = orc_encode_fixed(vy + P->dt * P->gravity, mult);
            c[2] = orc_encode_fixed(vz, mult);
            if (ox) c[0] = 0; if (oy) c[1] = 0; if (oz) c[2] = 0; /* X:479-481 */
        }
    }
}

static void update_range(void* ctx, int64_t lo, int64_t hi)
{
    StepCtx* s = (StepCtx*)ctx;
    for (int64_t i = lo; i < hi; ++i) update_cell(s->P, s->grid, i);
}
static void update_grid_impl(const OrcParams* P, void* grid, int nt)
{
    StepCtx s = {P, NULL, NULL, NULL, NULL, grid, 0};
    parallel_for(nt, num_cells(P), update_range, &s); /* F:382 */
}
void orc_update_grid(const OrcParams* P, void* grid) { update_grid_impl(P, grid, 1); }

/* ---------------------------------------------------------------- G2P  F:412-518, X:501-590, D:334-421, g2p.glsl */

static inline void cell_vel(const OrcParams* P, const void* grid, int64_t ci, float v[3])
{
    if (P->grid_mode == ORC_GRID_FLOAT) {
        const float* c = (const float*)grid + 4 * ci;
        v[0] = c[0]; v[1] = c[1]; v[2] = c[2];
    } else { /* X:531-535 */
        const int32_t* c = (const int32_t*)grid + 4 * ci;
        v[0] = orc_decode_fixed(c[0], P->fixed_point_mult);
        v[1] = orc_decode_fixed(c[1], P->fixed_point_mult);
        v[2] = orc_decode_fixed(c[2], P->fixed_point_mult);
    }
}

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void g2p_particle(const OrcParams* P, int64_t i, float* pos, float* vel, float* C,
                         const void* grid)
{
    const int Ry = P->grid[1], Rz = (P->dim == 3) ? P->grid[2] : 1;
    const int dim = P->dim;
    const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
    float wx[3], wy[3], wz[3];
    const int cx = axis_weights(px, wx), cy = axis_weights(py, wy);
    const int cz = (dim == 3) ? axis_weights(pz, wz) : 0;
    float v[3] = {0.0f, 0.0f, 0.0f};                             /* F:430 */
    float B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                    /* F:442  new Basis() is all-zero */
    if (dim == 2) {
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy) {
                float weight = wx[gx] * wy[gy];
                int nx = cx + gx - 1, ny = cy + gy - 1;
                int64_t ci = (int64_t)nx * Ry + ny;
                float dx = ((float)nx - px) + 0.5f, dy = ((float)ny - py) + 0.5f; /* D:361 */
                float gv[3];
                cell_vel(P, grid, ci, gv);
                float wvx = gv[0] * weight, wvy = gv[1] * weight; /* D:362 */
                B[0] += wvx * dx; B[1] += wvy * dx;               /* D:364-367 */
                B[3] += wvx * dy; B[4] += wvy * dy;
                v[0] += wvx; v[1] += wvy;                         /* D:369 */
            }
    } else {
        for (int gx = 0; gx < 3; ++gx)
            for (int gy = 0; gy < 3; ++gy)
                for (int gz = 0; gz < 3; ++gz) {
                    float weight = wx[gx] * wy[gy] * wz[gz];      /* F:449 */
                    int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1;
                    int64_t ci = ((int64_t)nx * Ry + ny) * Rz + nz;
                    float dx = ((float)nx - px) + 0.5f, dy = ((float)ny - py) + 0.5f,
                          dz = ((float)nz - pz) + 0.5f;           /* F:454 */
                    float gv[3];
                    cell_vel(P, grid, ci, gv);
                    float wvx = gv[0] * weight, wvy = gv[1] * weight, wvz = gv[2] * weight; /* F:455 */
                    B[0] += wvx * dx; B[1] += wvy * dx; B[2] += wvz * dx; /* F:457-462 */
                    B[3] += wvx * dy; B[4] += wvy * dy; B[5] += wvz * dy;
                    B[6] += wvx * dz; B[7] += wvy * dz; B[8] += wvz * dz;
                    v[0] += wvx; v[1] += wvy; v[2] += wvz;        /* F:464 */
                }
    }
    float* c = C + 9 * i;
    for (int k = 0; k < 9; ++k) c[k] = B[k] * 4.0f;              /* F:468-470 */

    /* advect + clamp  F:473-476 */
    float R[3] = {(float)P->grid[0], (float)P->grid[1], (float)P->grid[2]};
    float np_[3] = {px + v[0] * P->dt, py + v[1] * P->dt, pz + v[2] * P->dt};
    for (int a = 0; a < dim; ++a) np_[a] = clampf(np_[a], P->clamp_min, R[a] - P->clamp_max_off);
    if (dim == 2) np_[2] = pz;

    /* interaction */
    if (P->interaction == ORC_INTERACT_SPHERE_POST || P->interaction == ORC_INTERACT_SPHERE_PRE) {
        const float* q = (P->interaction == ORC_INTERACT_SPHERE_POST) ? np_ : &pos[3 * i];
        float dx = q[0] - P->sphere_pos[0], dy = q[1] - P->sphere_pos[1], dz = q[2] - P->sphere_pos[2];
        float d2 = (dx * dx + dy * dy) + dz * dz;                /* X:572 Dot */
        if (d2 < P->sphere_radius * P->sphere_radius) {
            float fx = 0, fy = 0, fz = 0;                        /* Normalized(): zero stays zero */
            if (d2 != 0) { float len = sqrtf(d2); fx = dx / len; fy = dy / len; fz = dz / len; }
            v[0] += fx * 1.0f; v[1] += fy * 1.0f; v[2] += fz * 1.0f; /* X:574-575 */
        }
        for (int k = 0; k < P->n_extra_spheres && k < 7; ++k) {  /* the same rule for every further sphere */
            const float* sp = P->extra_spheres[k];
            float ex = q[0] - sp[0], ey = q[1] - sp[1], ez = q[2] - sp[2];
            float e2 = (ex * ex + ey * ey) + ez * ez;
            if (e2 < sp[3] * sp[3]) {
                float gx = 0, gy = 0, gz = 0;
                if (e2 != 0) { float len = sqrtf(e2); gx = ex / len; gy = ey / len; gz = ez / len; }
                v[0] += gx * 1.0f; v[1] += gy * 1.0f; v[2] += gz * 1.0f;
            }
        }
    } else if (P->interaction == ORC_INTERACT_MOUSE_2D) {        /* D:384-397 */
        float dx = np_[0] - P->mouse_pos[0], dy = np_[1] - P->mouse_pos[1];
        float d2 = dx * dx + dy * dy;
        if (d2 < P->mouse_radius * P->mouse_radius) {
            float len = sqrtf(d2);
            float norm_factor = 1.0f / (len / P->mouse_radius);
            float nx = 0, ny = 0;
            if (d2 != 0) { nx = dx / len; ny = dy / len; }
            float fx = (nx * norm_factor) * 0.1f, fy = (ny * norm_factor) * 0.1f;
            if (!(isnan(fx) || isnan(fy))) { v[0] += fx; v[1] += fy; }
        }
    }

    /* predictive wall  F:506-514, D:409-416 */
    for (int a = 0; a < dim; ++a) {
        float xn = np_[a] + v[a];
        float wall_min = P->wall_min, wall_max = R[a] - P->wall_max_off;
        float va = v[a];
        if (xn < wall_min) va += P->wall_gain * (wall_min - xn); /* gain 1 is exact: F:509 */
        if (xn > wall_max) va += P->wall_gain * (wall_max - xn);
        v[a] = va;
    }
    pos[3 * i] = np_[0]; pos[3 * i + 1] = np_[1]; pos[3 * i + 2] = np_[2];
    vel[3 * i] = v[0]; vel[3 * i + 1] = v[1]; vel[3 * i + 2] = v[2];
}

static void g2p_range(void* ctx, int64_t lo, int64_t hi)
{
    StepCtx* s = (StepCtx*)ctx;
    for (int64_t i = lo; i < hi; ++i) g2p_particle(s->P, i, s->pos, s->vel, s->C, s->grid);
}
static void g2p_impl(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const void* grid,
                     int nt)
{
    StepCtx s = {P, pos, vel, C, NULL, (void*)grid, 0};
    parallel_for(nt, n, g2p_range, &s); /* F:419 */
}
void orc_g2p(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const void* grid)
{
    g2p_impl(P, n, pos, vel, C, grid, 1);
}

/* ---------------------------------------------------------------- Simulate  F:185-220 */

void orc_step(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const float* mass,
              void* grid, int32_t iterations)
{
    for (int it = 0; it < iterations; ++it) {
        clear_grid_impl(P, grid, 1);
        p2g1_impl(P, n, pos, vel, C, mass, grid, 1);
        p2g2_impl(P, n, pos, C, mass, grid, 1);
        update_grid_impl(P, grid, 1);
        g2p_impl(P, n, pos, vel, C, grid, 1);
    }
}

int32_t orc_step_mt(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const float* mass,
                    void* grid, int32_t iterations, int32_t nthreads)
{
    int nt = nthreads;
    if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN); /* MaxDegreeOfParallelism = -1  F:226 */
    if (nt < 1) nt = 1;
    for (int it = 0; it < iterations; ++it) {
        clear_grid_impl(P, grid, nt);
        p2g1_impl(P, n, pos, vel, C, mass, grid, nt);
        p2g2_impl(P, n, pos, C, mass, grid, nt);
        update_grid_impl(P, grid, nt);
        g2p_impl(P, n, pos, vel, C, grid, nt);
    }
    return nt;
}

/* ---------------------------------------------------------------- output hand-off  g2p.glsl:149-150 */

void orc_positions(int64_t n, const float* pos, const float* vel, float* out4)
{
    for (int64_t i = 0; i < n; ++i) {
        float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        out4[4 * i] = pos[3 * i]; out4[4 * i + 1] = pos[3 * i + 1]; out4[4 * i + 2] = pos[3 * i + 2];
        out4[4 * i + 3] = sqrtf((vx * vx + vy * vy) + vz * vz);
    }
}

/* ---------------------------------------------------------------- binning reference */

void orc_cell_keys(const OrcParams* P, int64_t n, const float* pos, int32_t* keys)
{
    const int Ry = P->grid[1], Rz = (P->dim == 3) ? P->grid[2] : 1;
    for (int64_t i = 0; i < n; ++i) {
        int cx = (int)pos[3 * i], cy = (int)pos[3 * i + 1], cz = (P->dim == 3) ? (int)pos[3 * i + 2] : 0;
        keys[i] = (cx * Ry + cy) * Rz + cz; /* F:259 + F:282 applied to the base cell */
    }
}

static int cmp_u64(const void* a, const void* b)
{
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

/* perm[j] = index of the particle that a stable sort by key puts at rank j */
void orc_stable_sort_perm(int64_t n, const uint32_t* keys, int32_t* perm)
{
    uint64_t* kv = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
    for (int64_t i = 0; i < n; ++i) kv[i] = ((uint64_t)keys[i] << 32) | (uint32_t)i;
    qsort(kv, (size_t)n, sizeof(uint64_t), cmp_u64); /* ties broken by original index = stable */
    for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)(kv[i] & 0xffffffffu);
    free(kv);
}
