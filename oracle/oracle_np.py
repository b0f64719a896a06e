"""Second, independent restatement of the reference MLS-MPM step in NumPy fp32 -- TEST INFRASTRUCTURE ONLY.

Purpose: the reference has no golden vectors (PARITY UNPINNED), so the C oracle
(mpm_oracle.c) is cross-checked against this vectorised restatement written from the same C# sources
(citation keys F / X / D / M / H as in mpm_oracle.h).  The two must agree bit-for-bit: integer grid words
in fixed-point mode, and float grid words too because np.add.at applies its updates sequentially in
flattened (particle-major, then gx, gy, gz) order, which is the reference's serial order (F:254-293).

Every array is float32 and every operator is one IEEE binary32 operation (NumPy never contracts to FMA).
"""
import numpy as np

f32 = np.float32
H = f32(0.5)


def _weights(p):
    """F:259-263.  p: (N,) float32 -> base cell (N,) int32 and weights (3, N)."""
    c = p.astype(np.int32)  # truncation toward zero
    cd = (p - c.astype(f32)) - H
    w0 = H * ((H - cd) * (H - cd))
    w1 = f32(0.75) - (cd * cd)
    w2 = H * ((H + cd) * (H + cd))
    return c, np.stack([w0, w1, w2])


def _stencil(P, pos):
    """Per particle, per stencil node (particle-major, gx, gy, gz order): weight, dist (3), cell index."""
    dim = P.dim
    Ry, Rz = P.grid[1], (P.grid[2] if dim == 3 else 1)
    cx, wx = _weights(pos[:, 0])
    cy, wy = _weights(pos[:, 1])
    if dim == 3:
        cz, wz = _weights(pos[:, 2])
        g = np.array([(a, b, c) for a in range(3) for b in range(3) for c in range(3)], np.int32)
    else:
        g = np.array([(a, b, 0) for a in range(3) for b in range(3)], np.int32)
    gx, gy, gz = g[:, 0], g[:, 1], g[:, 2]
    N = pos.shape[0]
    ar = np.arange(N)[:, None]
    weight = wx.T[ar, gx[None, :]] * wy.T[ar, gy[None, :]]      # F:273 (wx*wy)*wz
    nx = cx[:, None] + gx[None, :] - 1
    ny = cy[:, None] + gy[None, :] - 1
    dx = (nx.astype(f32) - pos[:, 0:1]) + H                     # F:276
    dy = (ny.astype(f32) - pos[:, 1:2]) + H
    if dim == 3:
        weight = weight * wz.T[ar, gz[None, :]]
        nz = cz[:, None] + gz[None, :] - 1
        dz = (nz.astype(f32) - pos[:, 2:3]) + H
        ci = (nx.astype(np.int64) * Ry + ny) * Rz + nz          # F:282
    else:
        dz = np.zeros_like(dx)
        ci = nx.astype(np.int64) * Ry + ny                      # D:224
    return weight.astype(f32), dx.astype(f32), dy.astype(f32), dz.astype(f32), ci


def encode(x, mult):
    return (x * f32(mult)).astype(np.int32)   # X:151-154 (values are in range: truncation)


def decode(i, mult):
    return i.astype(f32) / f32(mult)          # X:156-159


def _scatter(P, grid, ci, chans):
    """chans: dict column -> (N, S) float32 contributions."""
    flat = ci.reshape(-1)
    if P.grid_mode == 0:
        gf = grid.view(f32)
        for col, val in chans.items():
            np.add.at(gf[:, col], flat, val.reshape(-1).astype(f32))
    else:
        for col, val in chans.items():
            np.add.at(grid[:, col], flat, encode(val.reshape(-1), P.fixed_point_mult))


def clear_grid(P, grid):
    grid[...] = 0


def p2g1(P, pos, vel, C, mass, grid):
    w, dx, dy, dz, ci = _stencil(P, pos)
    c = [C[:, k:k + 1] for k in range(9)]
    if P.dim == 3:   # Basis * Vector3, row-dot  F:277
        qx = (c[0] * dx + c[3] * dy) + c[6] * dz
        qy = (c[1] * dx + c[4] * dy) + c[7] * dz
        qz = (c[2] * dx + c[5] * dy) + c[8] * dz
    else:            # Transform2D * Vector2 (+ zero origin)  D:229
        qx = (c[0] * dx + c[3] * dy) + f32(0)
        qy = (c[1] * dx + c[4] * dy) + f32(0)
        qz = np.zeros_like(qx)
    mc = w * mass[:, None]                                       # F:279
    ch = {3: mc, 0: mc * (vel[:, 0:1] + qx), 1: mc * (vel[:, 1:2] + qy)}
    if P.dim == 3:
        ch[2] = mc * (vel[:, 2:3] + qz)
    _scatter(P, grid, ci, ch)


def _pow(P, x):
    y = float(P.eos_power)
    if P.pow_mode == 1:
        raise NotImplementedError("libm powf mode exists only in the C oracle")
    if y == int(y) and 1 <= y <= 64:
        xd = x.astype(np.float64)
        r = xd.copy()
        for _ in range(int(y) - 1):
            r = r * xd
        return r.astype(f32)
    return np.power(x.astype(np.float64), np.float64(f32(P.eos_power))).astype(f32)


def p2g2(P, pos, C, mass, grid):
    w, dx, dy, dz, ci = _stencil(P, pos)
    if P.grid_mode == 0:
        gm = grid.view(f32)[:, 3][ci]
    else:
        gm = decode(grid[:, 3][ci], P.fixed_point_mult)          # X:383
    density = np.zeros(pos.shape[0], f32)
    for s in range(w.shape[1]):                                   # sequential sum, F:312-324
        density = density + gm[:, s] * w[:, s]
    volume = mass / density                                       # F:326
    pw = _pow(P, density / f32(P.rest_density))
    pr = f32(P.eos_stiffness) * (pw - f32(1))
    pressure = np.where(f32(-0.1) > pr, f32(-0.1), pr).astype(f32)  # F:331
    dt, visc = f32(P.dt), f32(P.dynamic_viscosity)
    c = [C[:, k] for k in range(9)]
    z = f32(0)
    if P.dim == 2:
        trace = c[3] + c[1]                                       # D:279
        t = {(0, 0): -pressure + visc * c[0], (0, 1): z + visc * trace,
             (1, 0): z + visc * trace, (1, 1): -pressure + visc * c[4]}   # (col,row)
        if P.eq16_order == 1:                                     # D:285
            s = (-dt) * volume
            e = {k: (s * v) * f32(4) for k, v in t.items()}
        else:                                                     # M:307
            s = (-volume) * f32(4)
            e = {k: (s * v) * dt for k, v in t.items()}
        momx = ((e[(0, 0)][:, None] * w) * dx + (e[(1, 0)][:, None] * w) * dy) + z
        momy = ((e[(0, 1)][:, None] * w) * dx + (e[(1, 1)][:, None] * w) * dy) + z
        _scatter(P, grid, ci, {0: momx, 1: momy})
        return
    # F:340-345 strain columns (col,row): strain.X = dudv.X + dudvT.X etc.
    sX = [c[0] + c[0], c[1] + c[3], c[2] + c[6]]
    sY = [c[3] + c[1], c[4] + c[4], c[5] + c[7]]
    sZ = [c[6] + c[2], c[7] + c[5], c[8] + c[8]]
    tX = [-pressure + sX[0] * visc, z + sX[1] * visc, z + sX[2] * visc]
    tY = [z + sY[0] * visc, -pressure + sY[1] * visc, z + sY[2] * visc]
    tZ = [z + sZ[0] * visc, z + sZ[1] * visc, -pressure + sZ[2] * visc]
    s = (-volume) * f32(4)                                        # F:347
    eX = [(s * v) * dt for v in tX]
    eY = [(s * v) * dt for v in tY]
    eZ = [(s * v) * dt for v in tZ]
    mom = []
    for r in range(3):                                            # F:364 row-dot
        mom.append(((eX[r][:, None] * w) * dx + (eY[r][:, None] * w) * dy) + (eZ[r][:, None] * w) * dz)
    _scatter(P, grid, ci, {0: mom[0], 1: mom[1], 2: mom[2]})


def update_grid(P, grid):
    dim = P.dim
    Rx, Ry, Rz = P.grid[0], P.grid[1], (P.grid[2] if dim == 3 else 1)
    G = grid.shape[0]
    i = np.arange(G)
    if dim == 3:
        x, y, zc = i // Rz // Ry, i // Rz % Ry, i % Rz           # F:399-401
    else:
        x, y, zc = i // Ry, i % Ry, np.full(G, 2)
    hi = P.bc_hi_off
    ox = (x < 2) | (x > Rx - hi)
    oy = (y < 2) | (y > Ry - hi)
    oz = ((zc < 2) | (zc > Rz - hi)) if dim == 3 else np.zeros(G, bool)
    dt, g = f32(P.dt), f32(P.gravity)
    if P.grid_mode == 0:
        gf = grid.view(f32)
        m = gf[:, 3]
        act = m > 0
        with np.errstate(all="ignore"):
            v = [gf[:, k] / m for k in range(3)]
        v[0] = v[0] + dt * f32(0)
        v[1] = v[1] + dt * g                                      # F:396
        v[2] = v[2] + dt * f32(0)
        if P.bc_mode == 0:
            v[0] = np.where(ox, f32(0), v[0]); v[1] = np.where(oy, f32(0), v[1]); v[2] = np.where(oz, f32(0), v[2])
        else:                                                     # M:366-368
            fr = f32(P.bc_friction)
            v1, v2 = np.where(ox, fr * v[1], v[1]), np.where(ox, fr * v[2], v[2]); v0 = np.where(ox, f32(0), v[0]); v = [v0, v1, v2]
            v0, v2 = np.where(oy, fr * v[0], v[0]), np.where(oy, fr * v[2], v[2]); v1 = np.where(oy, f32(0), v[1]); v = [v0, v1, v2]
            v0, v1 = np.where(oz, fr * v[0], v[0]), np.where(oz, fr * v[1], v[1]); v2 = np.where(oz, f32(0), v[2]); v = [v0, v1, v2]
        for k in range(3):
            gf[:, k] = np.where(act, v[k], gf[:, k]).astype(f32)
    else:
        mult = P.fixed_point_mult
        act = grid[:, 3] > 0
        mm = decode(grid[:, 3], mult)
        with np.errstate(all="ignore"):
            v = [decode(grid[:, k], mult) / mm for k in range(3)]
            v[1] = v[1] + dt * g                                  # X:471
            v = [np.where(act, a, f32(0)) for a in v]
            e = [encode(a, mult) for a in v]
        e[0] = np.where(ox, 0, e[0]); e[1] = np.where(oy, 0, e[1]); e[2] = np.where(oz, 0, e[2])
        for k in range(3):
            grid[:, k] = np.where(act, e[k], grid[:, k])


def g2p(P, pos, vel, C, grid):
    dim = P.dim
    w, dx, dy, dz, ci = _stencil(P, pos)
    if P.grid_mode == 0:
        gv = [grid.view(f32)[:, k][ci] for k in range(3)]
    else:
        gv = [decode(grid[:, k][ci], P.fixed_point_mult) for k in range(3)]
    N = pos.shape[0]
    v = [np.zeros(N, f32) for _ in range(3)]
    B = [np.zeros(N, f32) for _ in range(9)]
    d = [dx, dy, dz]
    for s in range(w.shape[1]):                                   # sequential over nodes F:443-466
        wv = [gv[k][:, s] * w[:, s] for k in range(3)]
        for col in range(dim):
            for row in range(dim):
                B[3 * col + row] = B[3 * col + row] + wv[row] * d[col][:, s]
        for k in range(dim):
            v[k] = v[k] + wv[k]
    for k in range(9):
        C[:, k] = B[k] * f32(4)                                   # F:468-470
    dt = f32(P.dt)
    R = [f32(P.grid[0]), f32(P.grid[1]), f32(P.grid[2])]
    old = pos.copy()
    newp = [pos[:, a] + v[a] * dt for a in range(3)]
    for a in range(dim):
        lo, hi = f32(P.clamp_min), R[a] - f32(P.clamp_max_off)
        newp[a] = np.where(newp[a] < lo, lo, np.where(newp[a] > hi, hi, newp[a])).astype(f32)
    if dim == 2:
        newp[2] = pos[:, 2]
    if P.interaction in (1, 2):
        q = newp if P.interaction == 1 else [old[:, 0], old[:, 1], old[:, 2]]
        sx, sy, sz = (q[a] - f32(P.sphere_pos[a]) for a in range(3))
        d2 = (sx * sx + sy * sy) + sz * sz
        inside = d2 < f32(P.sphere_radius) * f32(P.sphere_radius)
        with np.errstate(all="ignore"):
            ln = np.sqrt(d2)
            f = [np.where(d2 != 0, a / ln, f32(0)) for a in (sx, sy, sz)]
        for a in range(3):
            v[a] = np.where(inside, v[a] + f[a] * f32(1), v[a]).astype(f32)
    elif P.interaction == 3:
        mx, my = newp[0] - f32(P.mouse_pos[0]), newp[1] - f32(P.mouse_pos[1])
        d2 = mx * mx + my * my
        inside = d2 < f32(P.mouse_radius) * f32(P.mouse_radius)
        with np.errstate(all="ignore"):
            ln = np.sqrt(d2)
            nf = f32(1) / (ln / f32(P.mouse_radius))
            n = [np.where(d2 != 0, a / ln, f32(0)) for a in (mx, my)]
            fo = [(n[a] * nf) * f32(0.1) for a in range(2)]
        ok = inside & ~(np.isnan(fo[0]) | np.isnan(fo[1]))
        for a in range(2):
            v[a] = np.where(ok, v[a] + fo[a], v[a]).astype(f32)
    for a in range(dim):                                          # F:506-514
        xn = newp[a] + v[a]
        wmin, wmax, gain = f32(P.wall_min), R[a] - f32(P.wall_max_off), f32(P.wall_gain)
        va = np.where(xn < wmin, v[a] + gain * (wmin - xn), v[a])
        va = np.where(xn > wmax, va + gain * (wmax - xn), va)
        v[a] = va.astype(f32)
    for a in range(3):
        pos[:, a] = newp[a]
        vel[:, a] = v[a]


def step(P, pos, vel, C, mass, grid, iterations=1):
    for _ in range(iterations):
        clear_grid(P, grid)
        p2g1(P, pos, vel, C, mass, grid)
        p2g2(P, pos, C, mass, grid)
        update_grid(P, grid)
        g2p(P, pos, vel, C, grid)
