"""oracle_port.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

A numpy restatement of the reference's per-step particle hot path, independent
of the CUDA code, exposing the same host-side interface as
flecsph_b200.capi.BodySystem so that parity tests can drive either.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Parity status: PINNED.  tests/test_oracle.py checks every function below
against the UNMODIFIED reference headers compiled behind the C ABI
(oracle/_ref/libflecsph_ref_{1,2,3}d.so, built by oracle/Makefile from
/root/reference) on seeded inputs, and against the golden vectors committed
under tests/golden/ (generated from that same reference build by
tests/golden/make_golden.py).  The reference's own tests hold no numeric
golden vectors for this path; its behavioural pins
(neighbour set == brute force for uniform h, mpisph/test/tree3d.cc:87-108;
self in neighbours, app/drivers/tree/main_driver.cc:108-113; root mass == sum m,
tree3d.cc:85; key round trip, tree_topology/test/filling_curves.cc:202-241)
are asserted in tests/test_oracle.py as well.  Restated: the constant-viscosity /
ideal-or-polytropic hydro path of the drivers with external forces and potentials
(eforce.h; tests/test_eforce.py), all four equations of state, the relaxation phase
(drag, short-range repulsion), the Cullen-Dehnen viscosity, FMM gravity (monopole dual tree walk) and the periodic
boundaries (WALL copies); one rank.

Each function cites the reference file:line it follows (paths relative to
/root/reference).  Bit-exact parts (range, keys, order, tree, node summaries,
neighbour sets) use strictly the reference's operation order in IEEE fp64;
physics sums are only required to agree to 1e-10 and are vectorised.
"""
import math
import numpy as np

TINY = 1e10 * np.finfo(np.float64).tiny  # kernels.h:64, viscosity.h:36
SUB_ENTITIES = 128                       # tree_topology.h:2121
TOL = np.finfo(np.float64).eps * 10.0    # tree_geometry.h:57

VEC = ("pos", "vel", "velhalf", "acc", "gacc")
SCAL = ("mass", "h", "rho", "P", "u", "e", "cs", "dudt", "dedt", "dt", "alpha", "vsig", "gpot", "adiabatic", "ye",
        "divv", "ddivvdt", "trigger", "xi", "tracess", "gradv")

# sigma tables, kernels.h:40-47
SIGMA = {
    0: [4. / 3., 40. / (7. * math.pi), 8. / math.pi],
    1: [1. / 40., 63. / (478. * math.pi), 81. / (359. * math.pi)],
    2: [1.25, 7. / math.pi, 21. / (2. * math.pi)],
    3: [1.5, 9. / math.pi, 495. / (32. * math.pi)],
    4: [55. / 32., 78. / (7. * math.pi), 1365. / (64. * math.pi)],
}


def kernel_width(k):
    """kernels::select, kernels.h:753-764."""
    return 3.0 if k in (1, 5, 6) else 2.0


def _clz64(x):
    """count leading zeros of uint64 array (x != 0)."""
    x = x.astype(np.uint64)
    n = np.zeros(x.shape, np.int64)
    for s in (32, 16, 8, 4, 2, 1):
        m = (x >> np.uint64(64 - s)) == 0
        n += np.where(m, s, 0)
        x = np.where(m, x << np.uint64(s), x)
    return n


class ExternalForce:
    """external_force::select / acceleration / potential (include/physics/eforce.h:330-433) with the terms of :46-330 and the
    constant / parabolic / mesa profiles of include/physics/density_profiles.h:69-221,355-396, vectorised over the bodies.
    The terms are added in the order of the keyword, starting from zero."""

    def __init__(self, par, dim):
        self.par, self.dim, self.terms = par, dim, []
        kw = par.external_force_type.decode().strip()
        if kw.lower() in ("", "none", "zero"):
            return
        for it in kw.split(","):
            low = it.lower()
            if low in ("spherical wall", "airfoil", "spherical density support", "gravity", "orbit", "poison"):
                self.terms.append(low)
            elif low.startswith("walls:"):
                self.terms += ["wall" + str("xyz".index(ch)) for ch in low[6:9]]
            else:
                raise ValueError("bad external_force_type: " + it)
        if "spherical density support" in self.terms:
            prof = par.density_profile.decode().strip().lower() or "constant"
            if prof not in ("constant", "parabolic", "mesa"):
                raise NotImplementedError("density_profile " + prof)
            self.prof = prof
            if prof == "mesa":
                q = par.mesa_rim_width
                self.mesa_rho0 = {1: .5 / (1. - q / 3.), 2: 1. / (np.pi * self._mesa_helper(1.)), 3: 1. / (4. * np.pi * self._mesa_helper(1.))}[dim]
            self.K0 = par.pressure_initial / par.rho_initial ** par.poly_gamma
            self.rho0 = float(self._rho(np.zeros(1))[0])

    def _mesa_helper(self, r):
        dr = self.par.mesa_rim_width
        r0 = 1. - dr
        if self.dim == 2:
            return r ** 2 * (1. - (.5 * r ** 2 + r0 ** 2) / dr ** 2) + (.5 * r0 ** 3 + 4. * r ** 3) * r0 / (3. * dr ** 2)
        return r0 ** 3 / 3. + (r ** 3 - r0 ** 3) / 3. * (1. - r0 ** 2 / dr ** 2) + r0 * (r ** 2 - r0 ** 2) * (r ** 2 + r0 ** 2) / (2. * dr ** 2) \
            - (r ** 5 - r0 ** 5) / (5 * dr ** 2)

    def _rho(self, r):
        D = self.dim
        if self.prof == "parabolic":
            return {1: 0.75, 2: 2. / np.pi, 3: 15. / (8. * np.pi)}[D] * (1. - r * r)
        if self.prof == "mesa":
            q = self.par.mesa_rim_width
            return np.where(r < 1. - q, self.mesa_rho0, np.where(r < .9999, self.mesa_rho0 * (1. - (r - (1. - q)) ** 2 / q ** 2), 0.0))
        return np.full(len(r), {1: 0.5, 2: 1.0 / np.pi, 3: 0.75 / np.pi}[D])

    def _drhodr(self, r):
        D = self.dim
        if self.prof == "parabolic":
            return {1: -1.5, 2: -4. / np.pi, 3: -15. / (4. * np.pi)}[D] * r
        if self.prof == "mesa":
            q = self.par.mesa_rim_width
            return np.where((r > 1. - q) & (r < .9999), -2. * self.mesa_rho0 * (r - (1. - q)) / q ** 2, 0.0)
        return np.zeros(len(r))

    def _box(self):
        return (.5 * self.par.box_length, .5 * self.par.box_width, .5 * self.par.box_height)

    def _airfoil(self, x):
        p = self.par
        al = p.airfoil_attack_angle * np.pi / 180.0
        x1, y1 = x[:, 0] - p.airfoil_anchor_x, x[:, 1] - p.airfoil_anchor_y
        xx, yy = x1 * np.cos(al) + y1 * np.sin(al), -x1 * np.sin(al) + y1 * np.cos(al)
        inside = (np.abs(yy) < 5.0 * p.airfoil_thickness) & (xx > -p.airfoil_size * 0.02) & (xx < p.airfoil_size * 1.02)
        with np.errstate(invalid="ignore"):
            upper = p.airfoil_thickness * xx * np.sqrt(p.airfoil_size ** 2 - xx * xx)
        camber = p.airfoil_camber * np.sin(np.pi * xx / 2.)
        phi = upper ** 2 - (yy - camber) ** 2 + 0.002
        return al, xx, yy, camber, phi, inside & (phi > 0.0)

    def acceleration(self, x):
        p, D, n = self.par, self.dim, len(x)
        a = np.zeros((n, D))
        pn, pa = p.extforce_wall_powerindex, p.extforce_wall_steepness
        r = np.sqrt((x * x).sum(1))

        def sphere_wall():
            t = np.zeros((n, D))
            out = r > p.sphere_radius
            with np.errstate(invalid="ignore", divide="ignore"):
                ar = pn * pa * np.power(np.where(out, r - p.sphere_radius, 0.0), pn - 1)
                t[out] = (-x / r[:, None] * ar[:, None])[out]
            return t
        for term in self.terms:
            t = np.zeros((n, D))
            if term == "spherical wall":
                t = sphere_wall()
            elif term == "airfoil":
                al, xx, yy, camber, phi, on = self._airfoil(x)
                with np.errstate(invalid="ignore"):
                    pw = pn * pa * np.power(np.where(on, phi, 1.0), pn - 1)
                a0 = pw * (2. * (yy - camber) * (-p.airfoil_camber * np.pi / 2. * np.cos(np.pi / 2. * xx))
                           - p.airfoil_thickness ** 2 * 2 * xx * (p.airfoil_size ** 2 - 2 * xx * xx))
                a1 = pw * 2. * (yy - camber)
                t[:, 0] = np.where(on, a0 * np.cos(al) - a1 * np.sin(al), 0.0)
                t[:, 1] = np.where(on, a0 * np.sin(al) + a1 * np.cos(al), 0.0)
            elif term == "spherical density support":
                q = r / p.sphere_radius
                rho = p.rho_initial / self.rho0 * self._rho(q)
                drho = p.rho_initial / (self.rho0 * p.sphere_radius) * self._drhodr(q)
                with np.errstate(invalid="ignore", divide="ignore"):
                    a_r = np.where(rho > 0, self.K0 * p.poly_gamma * np.power(np.where(rho > 0, rho, 1.0), p.poly_gamma - 2) * drho, 0.0)
                    t = np.where((q > 1e-12)[:, None], a_r[:, None] * x / r[:, None], 0.0)
                t = t + sphere_wall()
            elif term == "gravity":
                t[:, 1 if D > 1 else 0] = -p.gravity_acceleration_constant
            elif term == "orbit":
                m_t = p.mass_neutron_star + p.mass_white_dwarf
                d2 = (x[:, 0] - p.orbital_separation) ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2
                t1 = -p.gravitational_constant * p.mass_neutron_star / np.sqrt(d2 ** 3)
                t2 = p.gravitational_constant * m_t / p.orbital_separation ** 3
                t[:, 0] = t1 * (x[:, 0] - p.orbital_separation) + t2 * (x[:, 0] - p.orbital_separation * p.mass_neutron_star / m_t)
                t[:, 1] = t1 * x[:, 1] + t2 * x[:, 1]
                t[:, 2] = t1 * x[:, 2]
            elif term.startswith("wall"):
                i = int(term[4])
                b, xi = self._box()[i], x[:, i]
                t[:, i] = (np.where(xi < -b, np.power(np.where(xi < -b, -xi - b, 0.0), pn - 1), 0.0)
                           - np.where(xi > b, np.power(np.where(xi > b, xi - b, 0.0), pn - 1), 0.0)) * pn * pa
            else:
                continue   # poison: potential only
            a = a + t
        return a

    def potential(self, x):
        p, D, n = self.par, self.dim, len(x)
        phi = np.zeros(n)
        pn, pa = p.extforce_wall_powerindex, p.extforce_wall_steepness
        r = np.sqrt((x * x).sum(1))
        sphere = np.where(r > p.sphere_radius, pa * np.power(np.where(r > p.sphere_radius, r - p.sphere_radius, 0.0), pn), 0.0)
        for term in self.terms:
            if term == "spherical wall":
                t = sphere
            elif term == "airfoil":
                _, _, _, _, aux, on = self._airfoil(x)
                t = np.where(on, pa * np.power(np.where(on, aux, 1.0), pn), 0.0)
            elif term == "spherical density support":
                rho = p.rho_initial / self.rho0 * self._rho(r / p.sphere_radius)
                t = np.where(rho > 0, -self.K0 * p.poly_gamma * np.power(np.where(rho > 0, rho, 1.0), p.poly_gamma - 1.) / (p.poly_gamma - 1.), 0.0) + sphere
            elif term == "gravity":
                t = x[:, 1 if D > 1 else 0] * p.gravity_acceleration_constant
            elif term == "orbit":
                m_t = p.mass_neutron_star + p.mass_white_dwarf
                t1 = -p.gravitational_constant * p.mass_neutron_star / np.sqrt((x[:, 0] - p.orbital_separation) ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2)
                t2 = -0.5 * p.gravitational_constant * m_t / p.orbital_separation ** 3
                t = t1 + t2 * ((x[:, 0] - p.orbital_separation * p.mass_neutron_star / m_t) ** 2 + x[:, 1] ** 2)
            elif term.startswith("wall"):
                i = int(term[4])
                b, xi = self._box()[i], x[:, i]
                t = (np.where(xi < -b, np.power(np.where(xi < -b, -xi - b, 0.0), pn), 0.0)
                     + np.where(xi > b, np.power(np.where(xi > b, xi - b, 0.0), pn), 0.0)) * pa
            else:
                t = np.full(n, p.zero_potential_poison_value)
            phi = phi + t
        return phi


class PortSystem:
    backend = "oracle-port"

    def __init__(self, params, backend=None, device=0):
        self.par = params
        # the port restates the constant-viscosity / ideal-or-polytropic path (with external forces); the other physics is
        # checked against oracle/_ref only -- refuse rather than silently ignore a selector
        self.ef = ExternalForce(params, params.dim)
        self.dim = D = params.dim
        self.B = 63 // D  # filling_curve.h:45-47
        self.n = 0
        self.c = {}
        self.dt_ = params.initial_dt
        self.totaltime = 0.0
        self.iteration = 0
        self.dt_saved = 0.0
        self.totaltime_next = 0.0
        self.t_out = [0.0, 0.0, 0.0]
        self.nbr_off = None
        self.kw = kernel_width(params.sph_kernel)

    def close(self):
        pass

    # ---- storage ------------------------------------------------------
    def resize(self, n):
        D = self.dim
        self.n = n
        for k in VEC:
            self.c[k] = np.zeros((n, D))
        for k in SCAL:
            self.c[k] = np.zeros(n)
        self.c["id"] = np.arange(n, dtype=np.int64)
        self.c["key"] = np.zeros(n, np.uint64)
        self.c["type"] = np.zeros(n, np.int32)
        self.c["nneigh"] = np.zeros(n, np.int32)

    def nbodies(self):
        return self.n

    def load_bodies(self, fields):
        self.resize(len(fields["mass"]))
        for k, v in fields.items():
            self.c[k] = np.array(v, dtype=self.c[k].dtype).reshape(self.c[k].shape)

    def set(self, col, arr):
        from flecsph_b200.capi import COL_NAMES
        k = COL_NAMES[col]
        self.c[k] = np.array(arr, dtype=self.c[k].dtype).reshape(self.c[k].shape)

    def get(self, col):
        from flecsph_b200.capi import COL_NAMES
        return self.c[COL_NAMES[col]].copy()

    def snapshot(self, names=None):
        names = names or [k for k in self.c if k != "nneigh"]
        return {k: self.c[k].copy() for k in names}

    def set_time(self, dt, totaltime=0.0, iteration=0):
        self.dt_, self.totaltime, self.iteration = dt, totaltime, iteration

    def get_time(self):
        return self.dt_, self.totaltime, self.iteration

    def advance_time(self):
        """physics::advance_time, default_physics.h:853-860."""
        self.iteration += 1
        if self.par.adaptive_timestep:
            self.totaltime = self.totaltime_next
        else:
            self.totaltime = self.totaltime + self.dt_

    # ---- range, keys, sort ---------------------------------------------
    def compute_range(self):
        """body_system::mpi_compute_range, mpisph/bodies_system.h:231-268."""
        x, h = self.c["pos"], self.c["h"]
        self.lo = (x - h[:, None]).min(0)
        self.hi = (x + h[:, None]).max(0)

    def get_range(self):
        return self.lo.copy(), self.hi.copy()

    def compute_keys(self):
        """morton_curve_u(range, p), filling_curve.h:503-525; compute_keys, tree_topology.h:828-832."""
        D, B = self.dim, self.B
        x = self.c["pos"]
        scale = self.hi - self.lo
        two_b = float(1 << B)
        t = (x - self.lo[None, :]) / scale[None, :] * two_b   # sub, div, mul in this order
        maxv = (1 << B) - 1
        key = np.full(self.n, 1 << (D * B), dtype=np.uint64)
        coords = []
        for d in range(D):
            # static_cast<uint64_t>(double): truncation; values reach 2^B at most
            cd = np.minimum(t[:, d], two_b)
            ci = np.array([min(maxv, int(v)) for v in cd], dtype=np.uint64) if B > 52 else \
                np.minimum(np.uint64(maxv), cd.astype(np.uint64))
            coords.append(ci)
        for i in range(B):
            for d in range(D):
                bit = (coords[d] >> np.uint64(i)) & np.uint64(1)
                key |= bit << np.uint64(i * D + d)
        self.c["key"] = key

    def sort_bodies(self):
        """psort::psort with the (key, id) comparator, bodies_system.h:195-206."""
        order = np.lexsort((self.c["id"], self.c["key"]))
        for k in self.c:
            self.c[k] = self.c[k][order]
        self.nbr_off = None

    # ---- tree -----------------------------------------------------------
    def build_tree(self):
        """tree_topology::build_tree, tree_topology.h:878-997 (result: the canonical octree of
        SURVEY.md 3.4) and physics::compute_cofm, default_physics.h:59-119."""
        D, B, n = self.dim, self.B, self.n
        key = self.c["key"]
        if n > 1:
            x = key[:-1] ^ key[1:]
            if (x == 0).any():
                raise RuntimeError("duplicate keys")
            p = 63 - _clz64(x)
            d = B - 1 - p // D
        else:
            d = np.zeros(0, np.int64)
        dprev = np.concatenate([[-1], d])
        dcur = np.concatenate([d, [-1]])
        if n == 1:
            dcur[0] = 0  # the root always exists (tree_topology.h:893)
        # cells in pre-order: nodes (dprev, dcur] then the leaf
        cd, cs, cl = [], [], []
        for i in range(n):
            for k in range(dprev[i] + 1, dcur[i] + 1):
                cd.append(k); cs.append(i); cl.append(0)
            cd.append(max(dprev[i], dcur[i]) + 1); cs.append(i); cl.append(1)
        cd, cs, cl = np.array(cd), np.array(cs), np.array(cl)
        order = np.lexsort((cs, cd))
        cd, cs, cl = cd[order], cs[order], cl[order]
        nc = len(cd)
        isnode = cl == 0
        noderank = np.cumsum(isnode) - isnode
        nn = int(isnode.sum())
        # parent prefix of each cell
        shift = (D * (B - (cd - 1))).astype(np.uint64)
        ppref = np.where(cd > 0, key[cs] >> np.minimum(shift, np.uint64(63)), np.uint64(0))
        head = np.zeros(nc, bool)
        head[1:] = (cd[1:] != cd[:-1]) | (ppref[1:] != ppref[:-1])
        head[0] = False
        H = np.cumsum(head)
        parent = H - 1  # node rank of the parent (root cell: -1)
        cell_of_node = np.nonzero(isnode)[0]
        child_begin = np.zeros(nn, np.int64)
        child_end = np.zeros(nn, np.int64)
        hb = np.nonzero(head)[0]
        child_begin[parent[hb]] = hb
        last = np.ones(nc, bool)
        last[:-1] = head[1:]
        last[0] = False
        le = np.nonzero(last)[0]
        child_end[parent[le]] = le + 1
        self.t = t = dict(cell_depth=cd, cell_start=cs, cell_isleaf=cl.astype(bool), cell_node=np.where(isnode, noderank, -1),
                          cell_parent=parent, child_begin=child_begin, child_end=child_end, nn=nn,
                          node_cell=cell_of_node, node_depth=cd[cell_of_node], node_start=cs[cell_of_node])
        t["node_key"] = key[t["node_start"]] >> (D * (B - t["node_depth"])).astype(np.uint64)
        self.leaf_depth_ = np.zeros(n, np.int32)
        lf = np.nonzero(cl == 1)[0]
        self.leaf_depth_[cs[lf]] = cd[lf]
        # upward pass, deepest nodes first (node ranks are in level order)
        pos, m, h = self.c["pos"], self.c["mass"], self.c["h"]
        mass = np.zeros(nn); com = np.zeros((nn, D)); radius = np.zeros(nn); lap = np.zeros(nn)
        bmin = np.zeros((nn, D)); bmax = np.zeros((nn, D)); sub = np.zeros(nn, np.int64)
        for nd in range(nn - 1, -1, -1):
            ch = range(child_begin[nd], child_end[nd])
            ents = [cs[c] for c in ch if cl[c] == 1]
            nods = [noderank[c] for c in ch if cl[c] == 0]
            X = np.zeros(D); M = 0.0; S = 0
            mn = np.full(D, np.finfo(np.float64).max); mx = -mn
            for e in ents:
                X = X + m[e] * pos[e]; M = M + m[e]; S += 1
                mn = np.minimum(mn, pos[e] - h[e] / 2.); mx = np.maximum(mx, pos[e] + h[e] / 2.)
            for c in nods:
                X = X + mass[c] * com[c]; M = M + mass[c]; S += sub[c]
                mn = np.minimum(mn, bmin[c]); mx = np.maximum(mx, bmax[c])
            X = X / M
            r = 0.0; lp = 0.0
            for e in ents:
                dist = self._distance(X, pos[e])
                r = max(r, dist); lp = max(lp, dist + h[e])
            for c in nods:
                dist = self._distance(X, com[c])
                r = max(r, dist + radius[c]); lp = max(lp, dist + radius[c] + lap[c])
            mass[nd] = M; com[nd] = X; radius[nd] = r; lap[nd] = lp; bmin[nd] = mn; bmax[nd] = mx; sub[nd] = S
        t.update(mass=mass, com=com, radius=radius, lap=lap, bmin=bmin, bmax=bmax, sub=sub)

    def gravitation_fmm(self):
        """body_system::gravitation_fmm (bodies_system.h:282-288): tree_topology::traversal_fmm (tree_topology.h:448-795) with the
        functors of include/physics/fmm.h (taylor_c2c / taylor_p2c :146-159, fmm_p2p :113-143, fmm_c2p :103-111) and
        fmm_sub_entities_ = 0.  A dual walk over (sink cell, source cell) pairs from (root, root): two bodies -> direct term;
        the same node -> all ordered pairs of its children; r1 + r2 < theta |X1 - X2| (tree_geometry.h:343-350) -> the source's
        monopole goes to the sink (a node accumulates it at its centre of mass in (pc, fc), a body takes it directly);
        otherwise the cell with the larger radius is split.  Afterwards every body adds pc - (x - X).fc and fc of each
        ancestor node that was a sink (interaction_c2p, fmm.h:80-97).  Accumulates into gpot / gacc (3-D only)."""
        assert self.dim == 3
        t, c, par = self.t, self.c, self.par
        cs, isleaf, cnode = t["cell_start"], t["cell_isleaf"], t["cell_node"]
        cb, ce = t["child_begin"], t["child_end"]
        mass, com, radius = t["mass"], t["com"], t["radius"]
        pos, m, ids = c["pos"], c["mass"], c["id"]
        gc, theta = par.gravitational_constant, par.fmm_macangle
        pc, fc, hit = np.zeros(t["nn"]), np.zeros((t["nn"], 3)), np.zeros(t["nn"], bool)
        gpot, gacc = c["gpot"].copy(), c["gacc"].copy()

        def monopole(x, X, M):
            d = self._distance(x, X)
            return -gc * M / d, -gc * M * (x - X) / (d * d * d)
        queue = [(0, 0)]
        while queue:
            nxt = []
            for c1, c2 in queue:
                l1, l2 = isleaf[c1], isleaf[c2]
                if l1 and l2:
                    i, j = cs[c1], cs[c2]
                    if ids[i] != ids[j]:
                        dp, da = monopole(pos[i], pos[j], m[j])
                        gpot[i] += dp
                        gacc[i] = gacc[i] + da
                    continue
                if c1 == c2:
                    kids = range(cb[cnode[c1]], ce[cnode[c1]])
                    for k1 in kids:
                        nxt.append((k1, k1))
                        for k2 in range(k1 + 1, kids.stop):
                            nxt.append((k1, k2))
                            nxt.append((k2, k1))
                    continue
                X1, r1 = (pos[cs[c1]], 0.0) if l1 else (com[cnode[c1]], radius[cnode[c1]])
                X2, r2, M2 = (pos[cs[c2]], 0.0, m[cs[c2]]) if l2 else (com[cnode[c2]], radius[cnode[c2]], mass[cnode[c2]])
                if r1 + r2 < theta * self._distance(X1, X2):
                    dp, da = monopole(X1, X2, M2)
                    if l1:
                        gpot[cs[c1]] += dp
                        gacc[cs[c1]] = gacc[cs[c1]] + da
                    else:
                        pc[cnode[c1]] += dp
                        fc[cnode[c1]] = fc[cnode[c1]] + da
                        hit[cnode[c1]] = True
                elif r1 > r2:
                    nxt += [(k, c2) for k in range(cb[cnode[c1]], ce[cnode[c1]])]
                else:
                    nxt += [(c1, k) for k in range(cb[cnode[c2]], ce[cnode[c2]])]
            queue = nxt
        # c2p: the bodies below every node that was a sink
        for nd in np.nonzero(hit)[0]:
            stack = [t["node_cell"][nd]]
            while stack:
                cell = stack.pop()
                if isleaf[cell]:
                    i = cs[cell]
                    r = pos[i] - com[nd]
                    gpot[i] += pc[nd] - (r * fc[nd]).sum()
                    gacc[i] = gacc[i] + fc[nd]
                else:
                    stack += list(range(cb[cnode[cell]], ce[cnode[cell]]))
        c["gpot"], c["gacc"] = gpot, gacc

    def _distance(self, a, b):
        """flecsi::distance, space_vector.h:46-58."""
        if self.dim == 1:
            return abs(a[0] - b[0])
        s = 0.0
        for d in range(self.dim):
            s = s + (a[d] - b[d]) * (a[d] - b[d])
        return math.sqrt(s)

    def tree_nodes(self):
        t = self.t
        o = np.argsort(t["node_key"], kind="stable")
        return dict(key=t["node_key"][o], mass=t["mass"][o], com=t["com"][o], radius=t["radius"][o], lap=t["lap"][o],
                    bmin=t["bmin"][o], bmax=t["bmax"][o], sub_entities=t["sub"][o])

    def leaf_depth(self):
        return self.leaf_depth_.copy()

    def periodic_walls(self):
        """boundary::pboundary_clean + pboundary_generate (include/physics/boundary.h:45-248) as update_iteration calls them
        (bodies_system.h:160-168): the WALL copies of the previous iteration go (stable), bodies outside the box re-enter on the
        other side, and every body within halo = 2.5 max h of a periodic face gets one copy per such axis plus the xy / xz / zy
        edge and the xyz corner copies, appended body by body in that order and numbered from the real-body count upwards."""
        c, p, D = self.c, self.par, self.dim
        keep = c["type"] != 1
        for k in list(c):
            c[k] = c[k][keep]
        self.n = n0 = int(keep.sum())
        halo = 2.5 * float(c["h"].max())
        periodic = (p.periodic_boundary_x, p.periodic_boundary_y, p.periodic_boundary_z)
        lo = np.array([-0.5 * p.box_length, -0.5 * p.box_width, -0.5 * p.box_height])[:D]
        hi = np.array([0.5 * p.box_length, 0.5 * p.box_width, 0.5 * p.box_height])[:D]
        pos = c["pos"]
        for d in range(D):
            if periodic[d]:
                x = pos[:, d]
                pos[:, d] = np.where(x > hi[d], x - (hi[d] - lo[d]), np.where(x <= lo[d], x + (hi[d] - lo[d]), x))
        src, newpos = [], []
        corners = {2: [(0, 1)], 3: [(0, 1), (0, 2), (2, 1), (0, 1, 2)]}.get(D, [])
        for i in range(n0):
            x = pos[i]
            up = [periodic[d] and bool(x[d] + halo > hi[d]) for d in range(D)]
            on = [periodic[d] and (up[d] or bool(x[d] - halo <= lo[d])) for d in range(D)]
            for d in range(D):
                if on[d]:
                    y = x.copy()
                    y[d] = y[d] - (hi[d] - lo[d]) if up[d] else y[d] + (hi[d] - lo[d])
                    src.append(i); newpos.append(y)
            for axes in corners:
                if all(on[d] for d in axes):
                    y = x.copy()
                    for d in axes:   # add_corner_XX / _XXX: the reference's own expressions
                        y[d] = lo[d] - abs(x[d] + lo[d]) if up[d] else hi[d] + abs(x[d] + hi[d])
                    src.append(i); newpos.append(y)
        if src:
            src = np.array(src)
            for k in list(c):
                c[k] = np.concatenate([c[k], c[k][src]])
            m = len(src)
            c["pos"][n0:] = np.array(newpos)
            c["type"][n0:] = 1
            c["id"][n0:] = n0 + np.arange(m)
            self.n = n0 + m

    def update_iteration(self):
        """body_system::update_iteration, bodies_system.h:152-229 (one rank)."""
        p = self.par
        if p.periodic_boundary_x or p.periodic_boundary_y or p.periodic_boundary_z:
            self.periodic_walls()
        self.compute_range()
        self.compute_keys()
        self.sort_bodies()
        self.build_tree()

    def reset_ghosts(self):
        """bodies_system.h:273 -> tree clean + rebuild (tree_topology.h:153-157): same tree."""
        self.build_tree()

    # ---- neighbour search -------------------------------------------------
    def work_cells(self):
        """traversal_sph work-cell collection, tree_topology.h:221-235."""
        t = self.t
        out = []
        stack = [0]
        while stack:
            c = stack.pop()
            if not t["cell_isleaf"][c]:
                nd = t["cell_node"][c]
                if t["sub"][nd] > SUB_ENTITIES:
                    stack.extend(range(t["child_begin"][nd], t["child_end"][nd]))
                    continue
            out.append(c)
        return out

    def _sphere_box(self, mn, mx, c, r):
        """tree_geometry::intersects_sphere_box: 3-D tree_geometry.h:329-341; 1-D :107-114; 2-D :219-227."""
        x = np.maximum(mn, np.minimum(c, mx))
        if self.dim == 3:
            d2 = (x[:, 0] - c[:, 0]) * (x[:, 0] - c[:, 0]) + (x[:, 1] - c[:, 1]) * (x[:, 1] - c[:, 1]) \
                + (x[:, 2] - c[:, 2]) * (x[:, 2] - c[:, 2])
            return d2 <= r * r
        if self.dim == 2:
            dist = np.sqrt((x[:, 0] - c[:, 0]) * (x[:, 0] - c[:, 0]) + (x[:, 1] - c[:, 1]) * (x[:, 1] - c[:, 1]))
        else:
            dist = np.abs(x[:, 0] - c[:, 0])
        return dist - r <= TOL

    def _dist2(self, a, b):
        """tree_geometry::distance2, tree_geometry.h:60,161,265."""
        s = (a[..., 0] - b[..., 0]) * (a[..., 0] - b[..., 0])
        for d in range(1, self.dim):
            s = s + (a[..., d] - b[..., d]) * (a[..., d] - b[..., d])
        return s

    def find_neighbors(self):
        """tree_topology::traversal_sph, tree_topology.h:212-442 (one rank): per work cell a
        breadth-first walk from the root with the three predicates of SURVEY.md A7."""
        if self.nbr_off is not None:
            return
        t = self.t
        pos, h = self.c["pos"], self.c["h"]
        lists = [None] * self.n
        for wc in self.work_cells():
            if t["cell_isleaf"][wc]:
                grp = np.array([t["cell_start"][wc]]); gnode = -1
            else:
                gnode = t["cell_node"][wc]
                s = t["node_start"][gnode]
                grp = np.arange(s, s + t["sub"][gnode])
            gp, gh = pos[grp], h[grp]
            cands = []
            queue = [0]
            while queue:
                nq = []
                for c in queue:
                    if not t["cell_isleaf"][c]:
                        nd = t["cell_node"][c]
                        if gnode >= 0:
                            # intersects_box_box, tree_geometry.h:306-314
                            if (t["bmax"][nd] < t["bmin"][gnode]).any() or (t["bmax"][gnode] < t["bmin"][nd]).any():
                                continue
                        if self._sphere_box(t["bmin"][nd][None, :], t["bmax"][nd][None, :], gp, gh).any():
                            nq.extend(range(t["child_begin"][nd], t["child_end"][nd]))
                    else:
                        e = t["cell_start"][c]
                        if gnode >= 0:
                            ext = max(h[e], t["lap"][gnode]) + t["radius"][gnode]
                            if not (self._dist2(pos[e], t["com"][gnode]) <= ext * ext):
                                continue
                        cands.append(e)
                queue = nq
            cands = np.array(cands, dtype=np.int64)
            cp, chh = pos[cands], h[cands]
            for k, g in enumerate(grp):
                r = np.maximum(gh[k], chh)
                ok = self._dist2(gp[k][None, :], cp) <= r * r
                lists[g] = np.sort(cands[ok])
        cnt = np.array([len(l) for l in lists])
        self.nbr_off = np.concatenate([[0], np.cumsum(cnt)])
        self.nbr_idx = np.concatenate(lists) if self.n else np.zeros(0, np.int64)
        self.nbr_a = np.repeat(np.arange(self.n), cnt)
        self.c["nneigh"] = cnt.astype(np.int32)

    def neighbor_lists(self):
        self.find_neighbors()
        ids = self.c["id"][self.nbr_idx]
        out = np.empty_like(ids)
        for i in range(self.n):
            s, e = self.nbr_off[i], self.nbr_off[i + 1]
            out[s:e] = np.sort(ids[s:e])
        return self.nbr_off.copy(), out

    # ---- SPH kernels (kernels.h) -------------------------------------------
    def W(self, r, h):
        D, k = self.dim, self.par.sph_kernel
        sg = SIGMA[k][D - 1]
        if k == 3:  # Wendland C4, kernels.h:375-410
            q = r / h
            if D == 1:
                return np.where(q < 1, sg / h * (1 - q) ** 5 * (1 + q * (5 + q * 8)), 0.0)
            return np.where(q < 1, sg / h ** D * (1 - q) ** 6 * (1 + q * (6 + q * 35 / 3)), 0.0)
        if k == 2:  # Wendland C2, kernels.h:282-320
            q = r / h
            if D == 1:
                return np.where(q < 1, sg / h * (1 - q) ** 3 * (1 + 3 * q), 0.0)
            return np.where(q < 1, sg / h ** D * (1 - q) ** 4 * (1 + 4 * q), 0.0)
        if k == 4:  # Wendland C6, kernels.h:480-520
            q = r / h
            if D == 1:
                return np.where(q < 1, sg / h * (1 - q) ** 7 * (1 + q * (7 + q * (19 + q * 21))), 0.0)
            return np.where(q < 1, sg / h ** D * (1 - q) ** 8 * (1 + q * (8 + q * (25 + q * 32))), 0.0)
        if k == 0:  # cubic spline, kernels.h:106-121
            q = 2. * r / h
            w = np.where(q < 1, 1 - 1.5 * q * q + .75 * q * q * q, .25 * (2 - q) ** 3)
            return np.where(q < 2, w * sg / h ** D, 0.0)
        if k == 1:  # quintic spline, kernels.h:217-234
            q = 3. * r / h
            w = (3 - q) ** 5 + np.where(q < 2, -6 * (2 - q) ** 5, 0.0) + np.where(q < 1, 15 * (1 - q) ** 5, 0.0)
            return np.where(q < 3, w * sg / h ** D, 0.0)
        raise NotImplementedError("kernel %d" % k)

    def gradW_over_r(self, r, h):
        """Scalar factor F such that grad W = r_ab * F (kernels.h gradients)."""
        D, k = self.dim, self.par.sph_kernel
        sg = SIGMA[k][D - 1]
        if k == 3:  # kernels.h:420-463
            q = r / h
            if D == 1:
                dW = -q * (1 - q) ** 4 * (1 + 4 * q)
            else:
                dW = -4. / 3. * q * (1 - q) ** 5 * (1 + 5 * q)
            return np.where(q < 1, 14. * sg / h ** (D + 1) * dW / (r + TINY), 0.0)
        if k == 2:  # kernels.h:330-365
            q = r / h
            if D == 1:
                return np.where(q < 1, 12. * sg / h ** 2 * (-q * (1 - q) ** 2) / (r + TINY), 0.0)
            return np.where(q < 1, 20. * sg / h ** (D + 1) * (-q * (1 - q) ** 3) / (r + TINY), 0.0)
        if k == 4:  # kernels.h:530-566
            q = r / h
            if D == 1:
                dW = -q * (1 - q) ** 6 * (3 + q * (18 + q * 35))
                return np.where(q < 1, 6. * sg / h ** 2 * dW / (r + TINY), 0.0)
            dW = -q * (1 - q) ** 7 * (1 + q * (7 + q * 16))
            return np.where(q < 1, 22. * sg / h ** (D + 1) * dW / (r + TINY), 0.0)
        if k == 0:  # kernels.h:133-153
            q = 2. * r / h
            dW = np.where(q < 1, -3.0 * q + 2.25 * q * q, -.75 * (2 - q) * (2 - q))
            return np.where(q < 2, 2. * sg / h ** (D + 1) * dW / (r + TINY), 0.0)
        if k == 1:  # kernels.h:246-268
            q = 3. * r / h
            dW = -5. * (3 - q) ** 4 + np.where(q < 2, 30. * (2 - q) ** 4, 0.0) + np.where(q < 1, -75. * (1 - q) ** 4, 0.0)
            return np.where(q < 3, 3. * sg / h ** (D + 1) * dW / (r + TINY), 0.0)
        raise NotImplementedError("kernel %d" % k)

    # ---- sweeps (default_physics.h) ------------------------------------------
    def _pairs(self):
        self.find_neighbors()
        a, b = self.nbr_a, self.nbr_idx
        c = self.c
        rab = c["pos"][a] - c["pos"][b]
        r = np.sqrt((rab * rab).sum(1)) if self.dim > 1 else np.abs(rab[:, 0])
        hab = .5 * (c["h"][a] + c["h"][b])
        return a, b, rab, r, hab

    def _sum(self, a, w):
        return np.bincount(a, weights=w, minlength=self.n)

    def _visc(self, a, b, rab, hab):
        """viscosity::mu, viscosity.h:81-93, viscosity_function<visc_constant> :128-137 and <visc_cullen> :150-157."""
        c, p = self.c, self.par
        v12 = c["velhalf"][a] - c["velhalf"][b]
        dot = (v12 * rab).sum(1)
        d2 = (rab * rab).sum(1)
        mu = hab * dot / (d2 + p.sph_viscosity_epsilon * hab * hab + TINY) * (dot < 0.0)
        rho_ab = .5 * (c["rho"][a] + c["rho"][b])
        c_ab = .5 * (c["cs"][a] + c["cs"][b])
        if p.sph_viscosity == 1:
            alpha_ab = .5 * (c["alpha"][a] + c["alpha"][b])
            return -alpha_ab * (c_ab - 2.0 * mu) * mu / rho_ab
        return (-p.sph_viscosity_alpha * c_ab + p.sph_viscosity_beta * mu) * mu / rho_ab

    def _eos(self, idx=slice(None)):
        """eos::compute_pressure + compute_soundspeed (eos/eos.h): ideal gas :166-183 from u, polytropic :100-118 from the
        body's adiabatic invariant K = P / rho^gamma, piecewise polytrope :365-394 (K = K1 of the first segment), cold
        white dwarf :254-275 from the electron fraction; eos_type none :425-432 leaves P and cs as they are."""
        c, p = self.c, self.par
        g = p.poly_gamma
        rho = c["rho"][idx]
        if p.eos_type == 0:
            c["P"][idx] = (g - 1.) * rho * c["u"][idx]
            c["cs"][idx] = np.sqrt(g * (g - 1.) * c["u"][idx])
        elif p.eos_type == 1:
            K = c["adiabatic"][idx]
            c["P"][idx] = K * np.power(rho, g)
            c["cs"][idx] = np.sqrt(K * g * np.power(rho, g - 1.))
        elif p.eos_type == 3:
            K, g2, thr = c["adiabatic"][idx], p.poly_gamma2, p.ppt_density_thr
            K2 = K * thr ** (g - g2)
            lo = rho < thr
            c["P"][idx] = np.where(lo, K * np.power(rho, g), K2 * np.power(rho, g2))
            c["cs"][idx] = np.where(lo, np.sqrt(K * g * np.power(rho, g - 1.)), np.sqrt(K2 * g2 * np.power(rho, g2 - 1.)))
        elif p.eos_type == 2:
            A_wd, B_wd, C_LIGHT = 6.00233181e22, 9.73932099e5, 2.99792458e10
            x = np.cbrt(rho * c["ye"][idx] / B_wd)
            x2 = x * x
            c["P"][idx] = A_wd * (x * (2. * x2 - 3.) * np.sqrt(x2 + 1.) + 3. * np.arcsinh(x))
            numer = (1. + x2) * (6. * x2 - 3.) + 3. + x2 * (2. * x2 - 3.)
            denom = (1. + x2) * (6. * x2 + 1.) - 1. + x2 * (2. * x2 + 1.)
            c["cs"][idx] = np.sqrt(numer / (3. * denom)) * C_LIGHT

    def _recover_u(self):
        """physics::recover_internal_energy, default_physics.h:126-145."""
        c = self.c
        eint = c["e"] - .5 * (c["vel"] * c["vel"]).sum(1) - self.ef.potential(c["pos"])
        if not (eint > 0).all():
            raise RuntimeError("internal energy non-positive")
        c["u"] = eint

    def apply_in_smoothinglength(self, op):
        c, p = self.c, self.par
        if op == 4:
            self.find_neighbors(); return
        a, b, rab, r, hab = self._pairs()
        if op == 0:
            # compute_density_pressure_soundspeed, default_physics.h:309-320
            rho = self._sum(a, c["mass"][b] * self.W(r, hab))
            if not (rho > 0).all():
                raise RuntimeError("density not positive")
            c["rho"] = rho
            cs_old = c["cs"].copy()
            if p.evolve_internal_energy and p.thermokinetic_formulation:
                self._recover_u()
            self._eos()
            # compute_signalspeed, default_physics.h:236-265; particles are processed in descending
            # key order, so b contributes its new cs iff index b > index a
            csb = np.where(b > a, c["cs"][b], cs_old[b])
            cab = np.maximum(c["cs"][a], csb)
            with np.errstate(invalid="ignore", divide="ignore"):
                n_ab = rab / r[:, None]
                vr = ((c["vel"][a] - c["vel"][b]) * n_ab).sum(1)
                val = cab - np.minimum(vr, 0.0)
            val = np.where(np.isnan(val), 0.0, val)
            vs = np.zeros(self.n)
            np.maximum.at(vs, a, val)
            c["vsig"] = vs
            if p.sph_viscosity == 1:
                # compute_divv, default_physics.h:273-300: div v = (v_a . sum_b m_b grad W_ab) / rho_a and its rate of change
                F = self.gradW_over_r(r, hab)
                dv = self._sum(a, c["mass"][b] * (c["vel"][a] * rab).sum(1) * F) / c["rho"]
                c["ddivvdt"] = (dv - c["divv"]) / self.dt_
                c["divv"] = dv
            return
        if op == 5:
            # viscosity::compute_alpha, viscosity.h:310-363 (A_trigger :283-308, compute_xi :193-272)
            vab = c["vel"][a] - c["vel"][b]
            cab = np.maximum(c["cs"][a], c["cs"][b])
            with np.errstate(invalid="ignore", divide="ignore"):
                val = cab - np.minimum((vab * (rab / r[:, None])).sum(1), 0.0)
            vs = np.zeros(self.n)
            np.maximum.at(vs, a, np.where(np.isnan(val), 0.0, val))    # the body itself (0/0) never wins the max
            F = self.gradW_over_r(r, hab)
            R = self._sum(a, np.sign(c["divv"][b]) * c["mass"][b] * self.W(r, hab)) / c["rho"]
            D = self.dim
            g = np.zeros((self.n, D, D))
            for i in range(D):
                for j in range(D):
                    g[:, i, j] = self._sum(a, c["mass"][b] * (-vab[:, i]) * rab[:, j] * F) / c["rho"]
            c["gradv"] = g[:, 0, 0].copy()
            sym = 0.5 * (g + g.transpose(0, 2, 1))
            for i in range(D):
                sym[:, i, i] -= c["divv"] / D
            tss = (sym * sym).sum((1, 2))
            c["tracess"] = tss
            res = (2.0 * (1.0 - R) ** 4 * c["divv"]) ** 2
            with np.errstate(invalid="ignore", divide="ignore"):
                xi = np.where(res < 1e-16, 0.0, res / (res + tss))
            trig = np.maximum(-c["ddivvdt"], 0.0)
            c["xi"], c["trigger"] = xi, trig
            At = xi * trig
            with np.errstate(invalid="ignore", divide="ignore"):
                aloc = p.sph_viscosity_alpha_max * At / (p.sph_viscosity_delta * (vs / c["h"]) ** 2 + At)
                decay = c["h"] / (2.0 * p.sph_viscosity_l * vs)
                damp = c["alpha"] * np.exp(-((c["alpha"] - aloc) / decay) * self.dt_)
            c["alpha"] = np.where(c["alpha"] <= aloc, aloc, damp)
            return
        if op == 6:
            # add_short_range_repulsion, default_physics.h:718-754: neighbours other than the body itself within
            # relaxation_repulsion_radius * h_ab
            take = (c["id"][a] != c["id"][b]) & ~(r > hab * p.relaxation_repulsion_radius)
            with np.errstate(invalid="ignore", divide="ignore"):
                w = np.where(take, c["mass"][b] / (r * r * r), 0.0)
            acc_r = np.stack([self._sum(a, np.where(take, w * rab[:, d], 0.0)) for d in range(self.dim)], 1)
            c["acc"] = c["acc"] + acc_r * p.relaxation_repulsion_gamma
            return
        F = self.gradW_over_r(r, hab)
        mt = c["mass"][b] * (rab != 0).any(1)
        Pi = self._visc(a, b, rab, hab)
        Pa = c["P"] / (c["rho"] * c["rho"])
        if op == 1:
            # compute_acceleration, default_physics.h:348-408
            coef = -mt * (Pa[a] + Pa[b] + Pi) * F
            acc = np.stack([self._sum(a, coef * rab[:, d]) for d in range(self.dim)], 1)
            acc = acc + self.ef.acceleration(c["pos"])     # default_physics.h:404
            c["acc"] = acc
            c["gacc"] = np.zeros_like(acc)
            c["gpot"] = np.zeros(self.n)
        elif op == 2:
            # compute_dudt, default_physics.h:421-487
            if self.iteration < p.relaxation_steps:
                c["dudt"] = np.zeros(self.n); return
            vdw = ((c["vel"][a] - c["vel"][b]) * rab).sum(1) * F
            c["dudt"] = Pa * self._sum(a, mt * vdw) + .5 * self._sum(a, mt * vdw * Pi)
        elif op == 3:
            # compute_dedt, default_physics.h:502-566
            va = (c["vel"][a] * rab).sum(1) * F
            vb = (c["vel"][b] * rab).sum(1) * F
            s = self._sum(a, mt * (Pa[a] * vb + va * Pa[b] + .5 * Pi * (vb + va)))
            c["dedt"] = -s + (c["gacc"] * c["vel"]).sum(1)
        else:
            raise ValueError(op)

    # ---- apply_all / get_all ---------------------------------------------------
    def apply_all(self, op):
        c, p, dt = self.c, self.par, self.dt_
        if op == 0:    # eos::init: adiabatic invariant (eos/eos.h:68-83,157-159; ppt :337-351), internal energy of the white dwarf :292-302
            if p.eos_type in (0, 1):
                c["adiabatic"] = c["P"] / np.power(c["rho"], p.poly_gamma)
            elif p.eos_type == 3:
                K2 = c["P"] / np.power(c["rho"], p.poly_gamma2)
                c["adiabatic"] = np.where(c["rho"] < p.ppt_density_thr, c["P"] / np.power(c["rho"], p.poly_gamma),
                                          K2 * p.ppt_density_thr ** (p.poly_gamma2 - p.poly_gamma))
            elif p.eos_type == 2:
                A_wd, B_wd = 6.00233181e22, 9.73932099e5
                x = np.cbrt(c["rho"] * c["ye"] / B_wd)
                x2, x3 = x * x, x * x * x
                c["u"] = A_wd / c["rho"] * (8. * x3 * (np.sqrt(x2 + 1.) - 1.) - (x * (2. * x2 - 3.) * np.sqrt(x2 + 1.) + 3. * np.arcsinh(x)))
        elif op == 1:  # set_total_energy, default_physics.h:327-335
            c["e"] = .5 * (c["vel"] * c["vel"]).sum(1) + c["u"] + self.ef.potential(c["pos"])
        elif op == 2:  # save_velocityhalf, integration.h:52
            c["velhalf"] = c["vel"].copy()
        elif op == 3:  # leapfrog_kick_v, integration.h:65
            c["vel"] = c["vel"] + 0.5 * dt * (c["acc"] + c["gacc"])
        elif op == 4:  # leapfrog_kick_u, integration.h:80
            c["u"] = c["u"] + 0.5 * dt * c["dudt"]
        elif op == 5:  # leapfrog_kick_e, integration.h:94
            c["e"] = c["e"] + 0.5 * dt * c["dedt"]
        elif op == 6:  # leapfrog_drift, integration.h:106
            c["pos"] = c["pos"] + dt * c["vel"]
        elif op == 7:  # recompute_pressure_soundspeed, default_physics.h:153-161
            u = c["u"].copy()
            c["u"] = u + 0.5 * dt * c["dudt"]
            self._eos()
            c["u"] = u
        elif op == 8:  # recompute_pressure_soundspeed_thermokinetic, default_physics.h:169-182
            self._recover_u()
            u = c["u"].copy()
            c["u"] = u + 0.5 * dt * (c["dedt"] - (c["vel"] * c["acc"]).sum(1))
            self._eos()
            c["u"] = u
        elif op == 9:  # compute_dt, default_physics.h:602-695
            tiny, mc = 1e-24, 0.6
            dx = c["h"] / (p.sph_eta * self.kw)
            vn = self._mag(c["vel"])
            dt_v = dx / (vn + tiny)
            acc = self._mag(c["acc"] + c["gacc"])
            dt_a = np.sqrt(dx / (acc + tiny))
            dt_c = dx / (tiny + c["vsig"] * (1 + mc * p.sph_viscosity_alpha))
            dtmin = p.timestep_cfl_factor * np.minimum(np.minimum(dt_v, dt_a), dt_c)
            if p.evolve_internal_energy and p.thermokinetic_formulation:
                if not (c["u"] > 0).all():
                    raise RuntimeError("internal energy non-positive")
                vg = (c["vel"] * c["gacc"]).sum(1)
                epot = self.ef.potential(c["pos"])
                done = np.zeros(self.n, bool)
                for _ in range(20):
                    depot = self.ef.potential(c["pos"] + dtmin[:, None] * c["vel"]) - epot if self.ef.terms else 0.0
                    done |= (depot + (-dtmin * vg)) < c["u"] * 0.5
                    dtmin = np.where(done, dtmin, dtmin * 0.25)
                if not done.all():
                    raise RuntimeError("eint-based dt estimator did not converge")
            c["dt"] = dtmin
        elif op == 12:  # viscosity::initialize_alpha, viscosity.h:172-180
            c["alpha"] = np.full(self.n, p.sph_viscosity_alpha if p.sph_viscosity == 0 else 0.0)
        elif op in (13, 14, 15):  # add_drag_acceleration :702-709, add_drag_dedt :573-581, add_drag_dudt :587-595
            v2 = (c["vel"] * c["vel"]).sum(1)
            drag = 0.0 - (p.relaxation_beta + p.relaxation_gamma * v2)[:, None] * c["vel"]     # external_force::acceleration_drag, eforce.h:441-451
            if op == 13:
                c["acc"] = c["acc"] + drag
            elif op == 14:
                c["dedt"] = c["dedt"] + (c["vel"] * drag).sum(1)
            else:
                c["dudt"] = c["dudt"] + (c["vel"] * drag).sum(1)
        elif op == 10:  # check_nans, default_physics.h:918-958
            for k in VEC + SCAL:
                if np.isnan(c[k]).any():
                    raise RuntimeError("NaN in " + k)
        elif op == 11:  # check_negativity, default_physics.h:962-979
            if (c["rho"] < 0).any() or (c["P"] < 0).any() or (p.evolve_internal_energy and (c["u"] < 0).any()):
                raise RuntimeError("negative quantity")
        else:
            raise ValueError(op)

    def _mag(self, v):
        return np.sqrt((v * v).sum(1)) if self.dim > 1 else np.abs(v[:, 0])

    def get_all(self, op, nout=3):
        c, p = self.c, self.par
        out = np.zeros(max(nout, 1))
        if op == 0:    # set_adaptive_timestep, default_physics.h:761-823
            dtmin = min(1e24, c["dt"].min())
            dt = self.dt_
            if dtmin < dt:
                dt = min(dtmin, dt / 2.0)
            if dtmin > 2.0 * dt:
                dt = dt * 2.0
            self.totaltime_next = self.totaltime + dt
            for j, odt in enumerate((p.out_screen_dt, p.out_scalar_dt, p.out_h5data_dt)):
                if odt > 0:
                    if self.totaltime == self.t_out[j]:
                        if self.dt_saved > 0:
                            dt = self.dt_saved
                        self.t_out[j] += odt
                    if self.totaltime + dt > self.t_out[j] - 0.1 * odt:
                        self.dt_saved = dt
                        self.totaltime_next = self.t_out[j] if j == 0 else min(self.totaltime_next, self.t_out[j])
            self.dt_ = self.totaltime_next - self.totaltime
            out[0] = self.dt_
        elif op in (1, 2):  # compute_smoothinglength, default_physics.h:825-848; average :879-898
            q = c["mass"] / c["rho"]
            r = q if self.dim == 1 else (np.sqrt(q) if self.dim == 2 else np.cbrt(q))
            c["h"] = r * p.sph_eta * self.kw
            if op == 2:
                c["h"] = np.full(self.n, 1. / self.n * c["h"].sum())
        elif op == 3:  # analysis::compute_total_mass, analysis.h:69-80 (NORMAL particles)
            out[0] = c["mass"][c["type"] == 0].sum()
        elif op == 4:  # analysis::compute_total_energy, analysis.h:113-150
            m = c["type"] == 0
            if p.thermokinetic_formulation:
                out[0] = (c["mass"] * c["e"])[m].sum()
            else:
                out[0] = (c["mass"] * (.5 * (c["vel"] * c["vel"]).sum(1) + c["u"] + self.ef.potential(c["pos"])))[m].sum()
            if p.enable_fmm:
                out[0] += .5 * (c["mass"] * c["gpot"])[m].sum()
        elif op == 5:  # analysis::compute_total_momentum, analysis.h:56-67
            m = c["type"] == 0
            out[:self.dim] = (c["mass"][:, None] * c["vel"])[m].sum(0)
        else:
            raise ValueError(op)
        return out

    def get_smoothinglength(self):
        """body_system::getSmoothinglength, bodies_system.h:112-127."""
        return float(self.c["h"].max())

    def step(self, initial=False):
        """One iteration of app/drivers/hydro/main_driver.cc:102-277 (+ newtonian/main_driver.cc:218-222 with enable_fmm)."""
        p = self.par
        S, A, G = self.apply_in_smoothinglength, self.apply_all, self.get_all

        relaxing = self.iteration < p.relaxation_steps     # hydro/main_driver.cc:128-133,152,165,205-206,221,235

        def acceleration():
            if p.sph_viscosity != 0:       # hydro/main_driver.cc:121-124,195-198
                S(5)
            self.reset_ghosts()
            S(1)
            if relaxing:
                A(13); S(6)
            if p.enable_fmm:               # newtonian/main_driver.cc:218-222
                self.gravitation_fmm()

        def energy():
            for m in range(1, p.pressure_updates_number + 1):
                if p.thermokinetic_formulation:
                    S(3)
                    if relaxing:
                        A(14)
                    A(8)
                else:
                    S(2)
                    if relaxing:
                        A(15)
                    A(7)
                if m < p.pressure_updates_number:
                    self.reset_ghosts()
        if initial:
            self.update_iteration()
            A(0)
            if p.thermokinetic_formulation:
                A(1)
            if p.sph_viscosity != 0:
                A(12)
            S(0); A(2); acceleration()
            if p.adaptive_timestep:
                A(9); G(0)
            if p.evolve_internal_energy:
                energy()
        else:
            if p.evolve_internal_energy:
                A(5 if p.thermokinetic_formulation else 4)
            A(3); A(2); A(6)
            self.update_iteration()
            S(0); acceleration()
            A(3); self.reset_ghosts()
            if p.evolve_internal_energy:
                energy()
                A(5 if p.thermokinetic_formulation else 4)
        if p.sph_variable_h:
            G(1)
        elif p.sph_update_uniform_h:
            G(2)
        A(10); A(11)
        if p.adaptive_timestep:
            A(9); G(0)
        self.advance_time()
