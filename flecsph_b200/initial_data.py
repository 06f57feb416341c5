"""Synthetic initial data for the benchmark / parity configurations.

Restates (does not copy) what the reference's single-rank initial-data
generators produce, so that no HDF5 file is needed:
  * lattices: include/lattice.h:224-293 (rectangular / HCP / FCC rules, loops
    `for(p = min; p < max; p += d)` -> sequential accumulation = np.cumsum) and
    in_domain_{1,3}d (lattice.h:74-133, b_tol = 1e-6);
  * Sedov blast:  app/id_generators/sedov/main.cc:68-312 with data/sedov_nx20.par;
  * Noh implosion: app/id_generators/noh/main.cc:236-292 with data/noh_nx20.par;
  * Sod tube 1-D: app/id_generators/sodtube/main.cc with data/sodtube_t1_n100.par
    (equal-mass variant, test #1);
  * Kelvin-Helmholtz 3-D: app/id_generators/KH/main.cc with data/KH_3d.par.
The sphere set-ups and Kelvin-Helmholtz are bit-equal to the H5part files the reference's own generators write
(tests/test_generators.py).
Every function returns (fields, params_overrides): a dict of column-name ->
numpy array in generation order (ids 0..N-1) and the `.par` values of the run.
"""
import math
import numpy as np

B_TOL = 1e-6  # lattice.h: particle_lattice::b_tol


def _axis(lo, hi, step, offset=0.0):
    """Values of `for(p = lo + offset; p < hi; p += step)` with the loop's own rounding."""
    n = int(math.ceil((hi - lo - offset) / step)) + 2
    steps = np.full(n, step)
    steps[0] = lo + offset
    p = np.cumsum(steps)
    return p[p < hi]


def in_domain_3d(X, Y, Z, bmin, bmax, domain_type):
    """in_domain_3d (lattice.h:107-133), operation by operation."""
    dx = B_TOL * (bmax[0] - bmin[0])
    xmin, xmax = bmin[0] - dx, bmax[0] - dx
    ymin, ymax = bmin[1] - dx, bmax[1] - dx
    zmin, zmax = bmin[2] - dx, bmax[2] - dx
    m = (X > xmin) & (X < xmax) & (Y > ymin) & (Y < ymax) & (Z > zmin) & (Z < zmax)
    if domain_type == 1:
        r = 0.5 * min(min(xmax - xmin, ymax - ymin), zmax - zmin)
        x0, y0, z0 = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax), 0.5 * (zmin + zmax)
        m &= ((X - x0) * (X - x0) + (Y - y0) * (Y - y0)) + (Z - z0) * (Z - z0) < r * r
    return m


def lattice_axes(lattice_type, bmin, bmax, sep):
    """The coordinate values the loops of generator_lattice_3d (lattice.h:224-293) visit, as tables: every loop
    variable is accumulated sequentially from its own start (`p += d`), so there is one z table, one y table per
    stacking state zl and one x table per (zl, row parity yo).  Returns (zs, zl_of_layer, ys[zl], xs[zl][yo])."""
    xmin, ymin, zmin = bmin
    xmax, ymax, zmax = bmax
    dx, dy, dz = sep, sep * math.sqrt(3.0) / 2.0, sep * math.sqrt(2.0 / 3.0)
    if lattice_type == 0:
        zs = _axis(zmin, zmax, dx)
        return zs, np.zeros(len(zs), np.int32), [_axis(ymin, ymax, dx)], [[_axis(xmin, xmax, dx)] * 2]
    zs = _axis(zmin, zmax, dz)
    nstate = 2 if lattice_type == 1 else 4
    zl_of = np.zeros(len(zs), np.int32)
    zl = 0
    for iz in range(len(zs)):
        zl_of[iz] = zl
        zl = 1 - zl if lattice_type == 1 else (zl + 1) * (1 if zl < 3 else 0)
    ys = [_axis(ymin - z * dy / 3, ymax, dy) for z in range(nstate)]
    xs = [[_axis(xmin + (yo - z) * dx / 2.0, xmax, dx) for yo in (0, 1)] for z in range(nstate)]
    return zs, zl_of, ys, xs


def lattice_3d(lattice_type, domain_type, bmin, bmax, sep, slab=None):
    """generator_lattice_3d (lattice.h:224-293). Returns [n,3] positions in generation order.
    slab = (k, K): only the z layers of the k-th of K contiguous groups of layers (one group per rank of a
    multi-GPU run; the layers of all groups together are the full lattice)."""
    zs, zl_of, ys_t, xs_t = lattice_axes(lattice_type, bmin, bmax, sep)
    out = []
    lay_lo, lay_hi = (0, len(zs)) if slab is None else ((slab[0] * len(zs)) // slab[1], ((slab[0] + 1) * len(zs)) // slab[1])
    for iz in range(lay_lo, lay_hi):
        z, zl = zs[iz], int(zl_of[iz])
        ys = ys_t[zl]
        x0, x1 = xs_t[zl][0], xs_t[zl][1]
        L = max(len(x0), len(x1))
        pad = lambda a: np.concatenate([a, np.full(L - len(a), np.inf)])
        rows = np.stack([pad(x0), pad(x1)])
        X = rows[np.arange(len(ys)) % 2]
        Y = np.broadcast_to(ys[:, None], X.shape)
        Z = np.full(X.shape, z)
        m = np.isfinite(X) & in_domain_3d(X, Y, Z, bmin, bmax, domain_type)
        out.append(np.stack([X[m], Y[m], Z[m]], 1))
    return np.concatenate(out, 0) if out else np.zeros((0, 3))


def _sphere_common(kind, nx, rho0, P0, gamma, eta, kernel_width, R=1.0, lattice_type=2, slab=None, allsum=None):
    """slab / allsum: generate only this rank's layers; allsum(int) -> the sum over all ranks (global counts)."""
    sep = 2.0 * R / (nx - 1)
    pos = lattice_3d(lattice_type, 1, (-R,) * 3, (R,) * 3, sep, slab=slab)
    n_local = len(pos)
    n = allsum(n_local) if allsum else n_local
    m, rho_st, rho_eos, h, u, P = sphere_scalars(kind, n, rho0, P0, gamma, eta, kernel_width, R)
    ids = np.arange(n_local, dtype=np.int64)
    if slab is not None:
        ids += np.int64(slab[0]) << np.int64(40)  # unique across ranks without a prefix sum of the counts
    n = n_local
    f = dict(pos=pos, vel=np.zeros((n, 3)), mass=np.full(n, m), h=np.full(n, h), rho=np.full(n, rho_st),
             u=np.full(n, u), id=ids)
    return f, m, rho_eos, P


def sphere_scalars(kind, n, rho0, P0, gamma, eta, kernel_width, R=1.0):
    """Per-body scalars of the constant-density sphere set-ups with each generator's own expressions:
    (mass, stored density, density its pressure is computed with, smoothing length, internal energy, pressure).
      sedov (sedov/main.cc:126-133,271-307) and noh (noh/main.cc:106-112,251-290): the density goes through the
        renormalised profile, rho_initial / profile(0) * profile(r/R) -- one rounding away from rho_initial -- and h, u, P
        follow from it; noh then stores rho_initial itself (noh/main.cc:260-261);
      collapse (collapse/main.cc:94-113,183-190): rho_initial, pressure_initial and u = P / (rho (gamma - 1)) as they are."""
    if kind == "collapse":
        total_mass = rho0 * 4.0 / 3.0 * math.pi * (R * R * R)
        m = total_mass / n
        return m, rho0, rho0, eta * kernel_width * math.pow(m / rho0, 1.0 / 3), P0 / (rho0 * (gamma - 1.0)), P0
    prof0 = 0.75 / math.pi                      # density_profiles::rho_constant_density, 3-D (density_profiles.h:69-81)
    total_mass = rho0 * (R * R * R) / prof0     # "normalize mass such that central density is rho_initial"
    m = total_mass / n
    rho_a = rho0 / prof0 * prof0
    h = eta * kernel_width * math.pow(m / rho_a, 1.0 / 3)
    K0 = P0 / math.pow(rho0, gamma)
    u = K0 * math.pow(rho_a, gamma - 1) / (gamma - 1)
    return m, (rho0 if kind == "noh" else rho_a), rho_a, h, u, None


def blast_energy_per_mass(blast_energy, m, nblast):
    """sedov/main.cc:236-243,293: u_blast = E / mass_blast with the blast mass accumulated body by body."""
    return blast_energy / float(np.cumsum(np.full(nblast, m))[-1])


def sedov_3d(nx=20, eta=1.2, blast_energy=2e-7, blast_radius=0.06, P0=1e-7, gamma=1.4, slab=None, allsum=None):
    """Sedov blast wave, data/sedov_nx20.par semantics; nx scales the particle count (~0.74 nx^3)."""
    f, m, rho_eos, _ = _sphere_common("sedov", nx, 1.0, P0, gamma, eta, 2.0, slab=slab, allsum=allsum)
    r = np.sqrt((f["pos"] ** 2).sum(1))
    blast = r < blast_radius
    nblast = allsum(int(blast.sum())) if allsum else int(blast.sum())
    if nblast:
        f["u"] = f["u"] + np.where(blast, blast_energy_per_mass(blast_energy, m, nblast), 0.0)
    f["P"] = rho_eos * f["u"] * (gamma - 1)
    f["dt"] = np.full(len(r), 1e-4)
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=gamma, initial_dt=1e-4, sph_variable_h=1,
               adaptive_timestep=1, timestep_cfl_factor=0.01)
    return f, par


def noh_3d(nx=20, eta=1.2, P0=1e-6, gamma=1.6666667, infall=0.1):
    """Noh implosion, data/noh_nx20.par semantics (noh/main.cc:236-292): radial inflow v = -0.1 r/|r|."""
    f, m, rho_eos, _ = _sphere_common("noh", nx, 1.0, P0, gamma, eta, 2.0)
    r = np.sqrt((f["pos"] ** 2).sum(1))
    rr = np.where(r > 0, r, 1.0)
    f["vel"] = -infall * f["pos"] / rr[:, None] * (r > 0)[:, None]
    f["P"] = rho_eos * f["u"] * (gamma - 1)
    f["dt"] = np.full(len(r), 2e-3)
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=gamma, initial_dt=2e-3, sph_variable_h=1,
               adaptive_timestep=1, timestep_cfl_factor=0.25)
    return f, par


def collapse_3d(nx=10, eta=1.2, P0=1e-10, gamma=1.4):
    """Self-gravitating cold sphere, data/collapse_nx10.par semantics: quintic spline, du/dt path, FMM gravity."""
    f, m, _, P = _sphere_common("collapse", nx, 1.0, P0, gamma, eta, 3.0)
    f["P"] = np.full(len(f["mass"]), P)
    f["dt"] = np.full(len(f["mass"]), 5e-4)
    par = dict(sph_kernel="quintic spline", sph_eta=eta, poly_gamma=gamma, initial_dt=5e-4, sph_variable_h=0,
               adaptive_timestep=0, thermokinetic_formulation=0, enable_fmm=1, fmm_macangle=0.8,
               gravitational_constant=1.0)
    return f, par


def sodtube_1d(nx=100, eta=1.2, gamma=1.4, box_length=3.0):
    """Sod shock tube test #1, 1-D, equal-mass particles (data/sodtube_t1_n100.par)."""
    nx = ((nx - 1) // 3) * 3 + 1  # sodtube/main.cc:124
    sep = box_length / (nx - 1)
    rho1, rho2, P1, P2 = 1.0, 0.125, 1.0, 0.1
    m = rho1 * sep
    sep_lr = m / rho2
    tol = B_TOL
    # in_domain_1d (lattice.h:74-83): keep xmin - tol*L < x < xmax - tol*L
    c = _axis(-box_length / 6.0, box_length / 6.0 - tol * box_length / 3.0, sep)
    rgt = _axis(box_length / 6.0, box_length / 2.0 - tol * box_length / 3.0, sep_lr)
    lft = _axis(-box_length / 2.0, -box_length / 6.0 - tol, sep_lr)
    x = np.concatenate([c, rgt, lft])
    mid = np.concatenate([np.ones(len(c), bool), np.zeros(len(rgt) + len(lft), bool)])
    n = len(x)
    rho = np.where(mid, rho1, rho2)
    P = np.where(mid, P1, P2)
    u = P / ((gamma - 1) * rho)
    h = eta * 2.0 * m / rho
    f = dict(pos=x[:, None], vel=np.zeros((n, 1)), mass=np.full(n, m), h=h, rho=rho, P=P, u=u,
             id=np.arange(n, dtype=np.int64), dt=np.full(n, 1e-3))
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=gamma, initial_dt=1e-3, sph_variable_h=1,
               adaptive_timestep=1, timestep_cfl_factor=0.25)
    return f, par


def random_box(n, dim=3, seed=0, h_jitter=0.0, eta=1.2, vel_amp=0.1):
    """Uniform random particles in the unit box (the reference's own tree tests use this:
    mpisph/test/tree3d.cc:48-111), optionally with +-h_jitter relative scatter in h."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, dim))
    m = np.full(n, 1.0 / n)
    h0 = eta * 2.0 * (1.0 / n) ** (1.0 / dim)
    h = h0 * (1.0 + h_jitter * (2.0 * rng.random(n) - 1.0))
    vel = vel_amp * (2.0 * rng.random((n, dim)) - 1.0)
    f = dict(pos=pos, vel=vel, mass=m, h=h, rho=np.ones(n), u=1.0 + 0.1 * rng.random(n), P=np.full(n, 0.4),
             id=rng.permutation(n).astype(np.int64), dt=np.full(n, 1e-3))
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=1.4, initial_dt=1e-3, sph_variable_h=1,
               adaptive_timestep=1, timestep_cfl_factor=0.25)
    return f, par


def periodic_box(n, dim=3, seed=0, h_jitter=0.0, eta=1.2, vel_amp=0.1, box=(1.0, 1.0, 1.0), periodic=(1, 1, 1), stray=0.0):
    """Random bodies in a box centred on the origin with periodic faces (the WALL-copy path, boundary.h:45-248);
    `stray` > 0 puts a few bodies just outside the box so that the wrap-around (boundary.h:141-158) is exercised."""
    rng = np.random.default_rng(seed)
    ext = np.array(box[:dim])
    pos = (rng.random((n, dim)) - 0.5) * ext
    if stray > 0:
        k = max(n // 50, 1)
        pos[:k] += np.sign(pos[:k]) * (0.5 * ext - np.abs(pos[:k])) + np.sign(pos[:k]) * stray * rng.random((k, dim))
    vol = float(np.prod(ext))
    m = np.full(n, vol / n)
    h0 = eta * 2.0 * (vol / n) ** (1.0 / dim)
    h = h0 * (1.0 + h_jitter * (2.0 * rng.random(n) - 1.0))
    vel = vel_amp * (2.0 * rng.random((n, dim)) - 1.0)
    f = dict(pos=pos, vel=vel, mass=m, h=h, rho=np.ones(n), u=1.0 + 0.1 * rng.random(n), P=np.full(n, 0.4),
             id=np.arange(n, dtype=np.int64), dt=np.full(n, 1e-3))
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=1.4, initial_dt=1e-3, sph_variable_h=1,
               adaptive_timestep=1, timestep_cfl_factor=0.25, box_length=box[0], box_width=box[1], box_height=box[2],
               periodic_boundary_x=periodic[0], periodic_boundary_y=periodic[1] if dim > 1 else 0,
               periodic_boundary_z=periodic[2] if dim > 2 else 0)
    return f, par


def kh_blocks(nx, lattice_type=1, density_ratio=2.0, box=(1.0, 1.0, 1.0)):
    """set_derived_params of app/id_generators/KH/main.cc:58-226 (3-D): the three lattice blocks (middle, top, bottom) as
    (bmin, bmax, separation), and the lengths the stretching step needs."""
    L, W, H = box
    sep = L / nx
    sep_t = sep * math.cbrt(density_ratio)
    dx_m = dy_m = dY_m = dz_m = dZ_m = sep
    dx_t = dy_t = dY_t = dz_t = dZ_t = sep_t
    if lattice_type in (1, 2):
        dy_m *= math.sqrt(3.) / 2.; dY_m = 2. * dy_m
        dy_t *= math.sqrt(3.) / 2.; dY_t = 2. * dy_t
        dz_m *= math.sqrt(2. / 3.); dz_t *= math.sqrt(2. / 3.)
        dZ_m = (2. if lattice_type == 1 else 3.) * dz_m
        dZ_t = (2. if lattice_type == 1 else 3.) * dz_t
    w_m = math.floor(W / (3. * dY_m)) * dY_m
    gap = min(dY_m, dY_t) / 2.
    w_t = math.floor((W / 2. - w_m / 2. - gap) / dY_t) * dY_t
    Nx_t = int(math.floor(L / dx_t + 0.1))
    q = 2. if lattice_type == 0 else 4.
    mx = (-L / 2. + dx_m / q, L / 2.)
    tx = (-Nx_t * dx_t / 2. + dx_t / q, Nx_t * dx_t / 2.)
    h_m = int(math.floor(H / dZ_m)) * dZ_m
    h_t = int(math.floor(H / dZ_t)) * dZ_t
    mid = ((mx[0], -0.5 * w_m, -h_m / 2.), (mx[1], 0.5 * w_m + 0.01 * dy_m, h_m / 2.), sep)
    top = ((tx[0], 0.5 * w_m + gap, -h_t / 2.), (tx[1], 0.5 * w_m + gap + w_t, h_t / 2.), sep_t)
    bot = ((tx[0], -0.5 * w_m - gap - w_t, -h_t / 2.), (tx[1], -0.5 * w_m - gap + 0.01 * dy_t, h_t / 2.), sep_t)
    return dict(blocks=(mid, top, bot), sep=sep, dx_t=dx_t, w_m=w_m, gap=gap, h_m=h_m, h_t=h_t)


def kh_3d_reference(nx=30, lattice_type=1, eta=1.5, gamma=1.4, density_ratio=2.0, flow_velocity=1.0, rho_initial=1.0,
                    pressure_initial=1.0, A=0.025, lam=0.1666, box=(1.0, 1.0, 1.0), cfl=0.25, kernel_width=2.0):
    """Kelvin-Helmholtz initial data as app/id_generators/KH/main.cc builds it (data/KH_3d.par): an HCP (or FCC /
    rectangular) block of density rho * ratio moving at +v/2 between two coarser blocks of density rho moving at -v/2,
    equal masses; the outer blocks are stretched to fit the periodic box (main.cc:296-326), a velocity perturbation a la
    Price (2008) sits at the interfaces (:352-356).  Statement by statement; bit-equal to the generator's H5part file
    (tests/test_generators.py).  Returns (fields, par); the generator's suggested time step is par["initial_dt"]."""
    L, W, H = box
    g = kh_blocks(nx, lattice_type, density_ratio, box)
    parts = [lattice_3d(lattice_type, 0, bmin, bmax, sp) for bmin, bmax, sp in g["blocks"]]
    npm, npt = len(parts[0]), len(parts[1])
    pos = np.concatenate(parts, 0)
    n = len(pos)
    x, y, z = pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy()
    rho_t, rho_m = rho_initial, rho_initial * density_ratio
    pmass = rho_m * g["sep"] * g["sep"] * g["sep"]
    if lattice_type in (1, 2):
        pmass *= 1. / math.sqrt(2.)
    # stretch the top and bottom blocks to the box (main.cc:296-326)
    yx = math.floor(L / g["dx_t"] + 0.1) * g["dx_t"] / L
    yz = g["h_t"] / g["h_m"]
    y0 = g["w_m"] / 2. + g["gap"]
    o = slice(npm, n)
    y[o] = np.where(y[o] > 0, y0 + yx * yz * (y[o] - y0), -y0 + yx * yz * (y[o] + y0))
    x[o] = x[o] / yx
    z[o] = z[o] / yz
    ys = .5 * W / abs(y[npm + npt])
    y = y * ys
    pmass *= ys
    zs = .5 * H / abs(z[npm + npt])
    z = z * zs
    pmass *= zs
    mid = np.arange(n) < npm
    rho = np.where(mid, rho_m, rho_t)
    vel = np.zeros((n, 3))
    vel[:, 0] = np.where(mid, flow_velocity / 2.0, -flow_velocity / 2.0)
    up = np.nonzero((y < 0.25) & (y > 0.25 - 0.025))[0]
    dn = np.nonzero((y > -0.25) & (y < -0.25 + 0.025))[0]
    vel[up, 1] = [A * math.sin(-2 * math.pi * (xx + .5) / lam) for xx in x[up]]     # libm sin, as the generator's
    vel[dn, 1] = [A * math.sin(2 * math.pi * (xx + .5) / lam) for xx in x[dn]]
    P = np.full(n, pressure_initial)
    u = P / (gamma - 1.) / rho
    h = np.where(mid, eta * kernel_width * math.pow(pmass / rho_m, 1. / 3), eta * kernel_width * math.pow(pmass / rho_t, 1. / 3))
    cs = math.sqrt(gamma * max(pressure_initial / rho_m, pressure_initial / rho_t))
    dt0 = cfl * g["sep"] / max(cs, flow_velocity)
    f = dict(pos=np.stack([x, y, z], 1), vel=vel, mass=np.full(n, pmass), h=h, rho=rho, u=u, P=P, id=np.arange(n, dtype=np.int64),
             dt=np.full(n, dt0))
    par = dict(sph_kernel="wendland c4", sph_eta=eta, poly_gamma=gamma, initial_dt=dt0, sph_variable_h=1, adaptive_timestep=1,
               timestep_cfl_factor=cfl, box_length=L, box_width=W, box_height=H, periodic_boundary_x=1, periodic_boundary_y=1,
               periodic_boundary_z=1)
    return f, par


def sphere_on_device(system, kind, nx, slab=None, allsum=None, eta=1.2, R=1.0, **kw):
    """The FCC-sphere set-ups (sedov_3d / noh_3d / collapse_3d above) built directly in the columns of a device system
    (fsph_generate_lattice / fsph_count_within / fsph_id_fill): the same bodies bit for bit, without the host arrays.
    Returns (n_local, par)."""
    spec = {"sedov": dict(P0=1e-7, gamma=1.4, kw_=2.0, dt=1e-4, v=0.0, blast=(2e-7, 0.06)),
            "noh": dict(P0=1e-6, gamma=1.6666667, kw_=2.0, dt=2e-3, v=-0.1, blast=None),
            "collapse": dict(P0=1e-10, gamma=1.4, kw_=3.0, dt=5e-4, v=0.0, blast=None)}[kind]
    spec.update(kw)
    rho0, P0, gamma = 1.0, spec["P0"], spec["gamma"]
    sep = 2.0 * R / (nx - 1)
    slab = slab or (0, 1)
    n_local = system.generate_lattice(2, 1, (-R,) * 3, (R,) * 3, sep, slab=slab, id_offset=(slab[0] << 40) if slab[1] > 1 else 0)
    n = allsum(n_local) if allsum else n_local
    m, rho_st, rho_eos, h, u, P_const = sphere_scalars(kind, n, rho0, P0, gamma, eta, spec["kw_"], R)
    add, rb = 0.0, 0.0
    if spec["blast"]:
        e_blast, rb = spec["blast"]
        nb = system.count_within((0.0, 0.0, 0.0), rb)
        nb = allsum(nb) if allsum else nb
        add = blast_energy_per_mass(e_blast, m, nb) if nb else 0.0
    system.id_fill(m, h, rho_st, u, add, rb, gamma, spec["dt"], spec["v"], rho_eos=rho_eos, pressure=P_const)
    par = {"sedov": sedov_3d, "noh": noh_3d, "collapse": collapse_3d}[kind](nx=4)[1]
    return n_local, par


def nx_for_particles(target, lo=8, hi=2000):
    """Smallest lattice_nx whose FCC unit sphere holds >= target particles (count ~ (4pi/3)/(s^3/sqrt2))."""
    est = lambda nx: (4.0 * math.pi / 3.0) / ((2.0 / (nx - 1)) ** 3 / math.sqrt(2.0))
    while lo < hi:
        mid = (lo + hi) // 2
        if est(mid) >= target:
            hi = mid
        else:
            lo = mid + 1
    return lo
