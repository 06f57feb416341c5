"""ctypes binding of the C ABI in include/flecsph_b200.h.

The same binding drives the product library (libflecsph_b200.so, CUDA only)
and -- from tests/ and bench.py's cpu_baseline leg only -- the two checker
libraries under oracle/.  It mirrors `body_system<double,D>` of the reference
(mpisph/bodies_system.h:38-379): update_iteration / reset_ghosts /
apply_in_smoothinglength / apply_all / get_all / gravitation_fmm.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# enum fsph_kernel / fsph_eos / fsph_viscosity
KERNELS = {"cubic spline": 0, "quintic spline": 1, "wendland c2": 2, "wendland c4": 3,
           "wendland c6": 4, "gaussian": 5, "super gaussian": 6, "sinc": 7}
EOS_IDEAL, EOS_POLYTROPIC, EOS_WD, EOS_PPT, EOS_NONE = range(5)
VISC_CONSTANT, VISC_CULLEN = 0, 1

# enum fsph_column
(COL_POS, COL_VEL, COL_VELHALF, COL_ACC, COL_GACC, COL_MASS, COL_H, COL_RHO, COL_P, COL_U, COL_E,
 COL_CS, COL_DUDT, COL_DEDT, COL_DT, COL_ALPHA, COL_VSIG, COL_GPOT, COL_ADIABATIC, COL_ID, COL_KEY,
 COL_TYPE, COL_NNEIGH, COL_YE, COL_DIVV, COL_DDIVVDT, COL_TRIGGER, COL_XI, COL_TRACESS, COL_GRADV) = range(30)
VECTOR_COLS = (COL_POS, COL_VEL, COL_VELHALF, COL_ACC, COL_GACC)
COL_DTYPE = {COL_ID: np.int64, COL_KEY: np.uint64, COL_TYPE: np.int32, COL_NNEIGH: np.int32}
COL_NAMES = ["pos", "vel", "velhalf", "acc", "gacc", "mass", "h", "rho", "P", "u", "e", "cs", "dudt",
             "dedt", "dt", "alpha", "vsig", "gpot", "adiabatic", "id", "key", "type", "nneigh", "ye", "divv", "ddivvdt",
             "trigger", "xi", "tracess", "gradv"]

# enum fsph_sweep_op
(SWEEP_DENSITY, SWEEP_ACCELERATION, SWEEP_DUDT, SWEEP_DEDT, SWEEP_NEIGHBORS_ONLY, SWEEP_COMPUTE_ALPHA,
 SWEEP_SHORT_RANGE_REPULSION) = range(7)
# enum fsph_apply_op
(APPLY_EOS_INIT, APPLY_SET_TOTAL_ENERGY, APPLY_SAVE_VELOCITYHALF, APPLY_KICK_V, APPLY_KICK_U, APPLY_KICK_E,
 APPLY_DRIFT, APPLY_RECOMPUTE_P_CS, APPLY_RECOMPUTE_P_CS_TK, APPLY_COMPUTE_DT, APPLY_CHECK_NANS,
 APPLY_CHECK_NEGATIVITY, APPLY_INITIALIZE_ALPHA, APPLY_ADD_DRAG_ACCELERATION, APPLY_ADD_DRAG_DEDT, APPLY_ADD_DRAG_DUDT) = range(16)
# enum fsph_reduce_op
(GET_SET_ADAPTIVE_TIMESTEP, GET_COMPUTE_SMOOTHINGLENGTH, GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH, GET_TOTAL_MASS,
 GET_TOTAL_ENERGY, GET_TOTAL_MOMENTUM, GET_TOTAL_KINETIC_ENERGY, GET_TOTAL_INTERNAL_ENERGY,
 GET_TOTAL_GRAVITATIONAL_ENERGY, GET_TOTAL_ANG_MOM, GET_SMOOTHINGLENGTH_STATS, GET_VELOCITY_STATS,
 GET_NEIGHBORS_STATS) = range(13)


class Params(C.Structure):
    """struct fsph_params."""
    _fields_ = [
        ("dim", C.c_int32), ("sph_kernel", C.c_int32), ("eos_type", C.c_int32), ("sph_viscosity", C.c_int32),
        ("sph_eta", C.c_double), ("sph_sinc_index", C.c_double), ("poly_gamma", C.c_double),
        ("sph_viscosity_alpha", C.c_double), ("sph_viscosity_beta", C.c_double),
        ("sph_viscosity_epsilon", C.c_double), ("timestep_cfl_factor", C.c_double), ("initial_dt", C.c_double),
        ("out_screen_dt", C.c_double), ("out_scalar_dt", C.c_double), ("out_h5data_dt", C.c_double),
        ("fmm_macangle", C.c_double), ("gravitational_constant", C.c_double),
        ("box_length", C.c_double), ("box_width", C.c_double), ("box_height", C.c_double),
        ("adaptive_timestep", C.c_int32), ("sph_variable_h", C.c_int32), ("sph_update_uniform_h", C.c_int32),
        ("evolve_internal_energy", C.c_int32), ("thermokinetic_formulation", C.c_int32), ("enable_fmm", C.c_int32),
        ("relaxation_steps", C.c_int32), ("pressure_updates_number", C.c_int32),
        ("periodic_boundary_x", C.c_int32), ("periodic_boundary_y", C.c_int32), ("periodic_boundary_z", C.c_int32),
        ("reserved_", C.c_int32),
        ("sph_viscosity_alpha_max", C.c_double), ("sph_viscosity_l", C.c_double), ("sph_viscosity_delta", C.c_double),
        ("poly_gamma2", C.c_double), ("ppt_density_thr", C.c_double), ("relaxation_beta", C.c_double),
        ("relaxation_gamma", C.c_double), ("relaxation_repulsion_radius", C.c_double), ("relaxation_repulsion_gamma", C.c_double),
        # external forces (include/physics/eforce.h)
        ("external_force_type", C.c_char * 64), ("density_profile", C.c_char * 32),
        ("sphere_radius", C.c_double), ("extforce_wall_powerindex", C.c_double), ("extforce_wall_steepness", C.c_double),
        ("gravity_acceleration_constant", C.c_double), ("zero_potential_poison_value", C.c_double),
        ("mesa_rim_width", C.c_double), ("rho_initial", C.c_double), ("pressure_initial", C.c_double),
        ("orbital_separation", C.c_double), ("mass_neutron_star", C.c_double), ("mass_white_dwarf", C.c_double),
        ("airfoil_size", C.c_double), ("airfoil_thickness", C.c_double), ("airfoil_camber", C.c_double),
        ("airfoil_anchor_x", C.c_double), ("airfoil_anchor_y", C.c_double), ("airfoil_attack_angle", C.c_double),
    ]


class StreamOut(C.Structure):   # fsph_stream_out
    _fields_ = [("col", C.c_int32), ("comp", C.c_int32), ("host", C.c_void_p)]


# every symbol include/flecsph_b200.h declares
SYMBOLS = [
    "fsph_default_params", "fsph_create", "fsph_destroy", "fsph_last_error", "fsph_backend", "fsph_resize",
    "fsph_nbodies", "fsph_set_column", "fsph_get_column", "fsph_set_component", "fsph_get_component", "fsph_launch_count", "fsph_upload_aos", "fsph_download_aos", "fsph_set_time",
    "fsph_get_time", "fsph_advance_time", "fsph_update_iteration", "fsph_reset_ghosts", "fsph_sweep", "fsph_apply",
    "fsph_reduce", "fsph_gravitation_fmm", "fsph_get_range", "fsph_get_smoothinglength", "fsph_step", "fsph_step_streamed",
    "fsph_step_pipelined", "fsph_stream_wait",
    "fsph_tree_num_nodes", "fsph_tree_get_nodes", "fsph_tree_get_leaf_depth", "fsph_neighbor_lists",
    "fsph_phase_times", "fsph_set_time_state", "fsph_get_time_state", "fsph_tree_root",
    "fsph_comm_unique_id", "fsph_comm_init_nccl", "fsph_local_group_create", "fsph_local_group_destroy",
    "fsph_comm_init_local", "fsph_comm_rank", "fsph_comm_size", "fsph_nghosts", "fsph_dd_choose_splitters",
    "fsph_dd_align_splitter", "fsph_measure_fp64_tflops", "fsph_measure_copy_gbs",
    "fsph_generate_lattice", "fsph_count_within", "fsph_id_fill",
]
# entry points only the product library has to provide (the checkers under oracle/ may omit them)
PRODUCT_ONLY = {"fsph_step_streamed", "fsph_step_pipelined", "fsph_stream_wait", "fsph_set_time_state", "fsph_get_time_state", "fsph_tree_root", "fsph_comm_unique_id",
                "fsph_comm_init_nccl", "fsph_local_group_create", "fsph_local_group_destroy", "fsph_comm_init_local",
                "fsph_comm_rank", "fsph_comm_size", "fsph_nghosts", "fsph_dd_choose_splitters", "fsph_dd_align_splitter",
                "fsph_measure_fp64_tflops", "fsph_measure_copy_gbs", "fsph_generate_lattice", "fsph_count_within",
                "fsph_id_fill"}

# FSPH_PRODUCT_LIB: another build of the same library (kernel-variant A/B runs of scripts/ab_bench.sh)
PRODUCT_LIB = os.environ.get("FSPH_PRODUCT_LIB") or os.path.join(ROOT, "flecsph_b200", "libflecsph_b200.so")


def library_path(which, dim=3):
    """'cuda' -> the product; 'ref' / 'ref_fast' -> the reference headers behind the ABI under oracle/_ref (tests and bench
    baseline only; the numpy restatement oracle/oracle_port.py is imported directly by the tests)."""
    if which == "cuda":
        return PRODUCT_LIB
    if which == "ref":
        return os.path.join(ROOT, "oracle", "_ref", "libflecsph_ref_%dd.so" % dim)
    if which == "ref_fast":
        return os.path.join(ROOT, "oracle", "_ref", "libflecsph_ref_3d_fast.so")
    raise ValueError(which)


_LIBS = {}


def load(path):
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no fallback path)" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "fsph_default_params": (None, [C.POINTER(Params), i32]),
        "fsph_create": (vp, [C.POINTER(Params), i32]),
        "fsph_destroy": (None, [vp]),
        "fsph_last_error": (C.c_char_p, [vp]),
        "fsph_backend": (C.c_char_p, []),
        "fsph_resize": (i32, [vp, i64]),
        "fsph_nbodies": (i64, [vp]),
        "fsph_set_column": (i32, [vp, i32, vp]),
        "fsph_get_column": (i32, [vp, i32, vp]),
        "fsph_set_component": (i32, [vp, i32, i32, vp]),
        "fsph_get_component": (i32, [vp, i32, i32, vp]),
        "fsph_launch_count": (C.c_uint64, []),
        "fsph_upload_aos": (i32, [vp, vp, i64, i64, vp]),
        "fsph_download_aos": (i32, [vp, vp, i64, vp]),
        "fsph_set_time": (i32, [vp, dbl, dbl, i64]),
        "fsph_get_time": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)]),
        "fsph_advance_time": (i32, [vp]),
        "fsph_update_iteration": (i32, [vp]),
        "fsph_reset_ghosts": (i32, [vp]),
        "fsph_sweep": (i32, [vp, i32]),
        "fsph_apply": (i32, [vp, i32]),
        "fsph_reduce": (i32, [vp, i32, vp, i32]),
        "fsph_gravitation_fmm": (i32, [vp]),
        "fsph_get_range": (i32, [vp, vp, vp]),
        "fsph_get_smoothinglength": (dbl, [vp]),
        "fsph_step": (i32, [vp, i32]),
        "fsph_step_streamed": (i32, [vp, i32, vp, i32]),
        "fsph_step_pipelined": (i32, [vp, i32, vp, i32, vp, i32]),
        "fsph_stream_wait": (i32, [vp]),
        "fsph_tree_num_nodes": (i64, [vp]),
        "fsph_tree_get_nodes": (i32, [vp] + [vp] * 8),
        "fsph_tree_get_leaf_depth": (i32, [vp, vp]),
        "fsph_neighbor_lists": (i32, [vp, vp, vp]),
        "fsph_phase_times": (i32, [vp, vp, i32, C.POINTER(C.c_char_p)]),
        "fsph_set_time_state": (i32, [vp, vp, i64]),
        "fsph_get_time_state": (i32, [vp, vp, C.POINTER(i64)]),
        "fsph_tree_root": (i32, [vp, vp, C.POINTER(dbl)]),
        "fsph_comm_unique_id": (i32, [vp]),
        "fsph_comm_init_nccl": (i32, [vp, vp, i32, i32]),
        "fsph_local_group_create": (vp, [i32]),
        "fsph_local_group_destroy": (None, [vp]),
        "fsph_comm_init_local": (i32, [vp, vp, i32]),
        "fsph_comm_rank": (i32, [vp]),
        "fsph_comm_size": (i32, [vp]),
        "fsph_nghosts": (i64, [vp]),
        "fsph_dd_choose_splitters": (None, [i32, i32, vp, vp, vp]),
        "fsph_dd_align_splitter": (C.c_uint64, [i32, vp, i32, vp, i32, C.c_uint64]),
        "fsph_measure_fp64_tflops": (dbl, [i32, i32]),
        "fsph_measure_copy_gbs": (dbl, [i32, i64, i32]),
        "fsph_generate_lattice": (i32, [vp, i32, i32, vp, vp, dbl, i32, i32, i64, C.POINTER(i64)]),
        "fsph_count_within": (i32, [vp, vp, dbl, C.POINTER(i64)]),
        "fsph_id_fill": (i32, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        if name in PRODUCT_ONLY and not hasattr(lib, name):
            continue
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _LIBS[path] = lib
    return lib


def default_params(dim=3, lib=None, **kw):
    """Reference defaults (include/params.h) with overrides; kernel may be given by its .par name."""
    p = Params()
    if lib is not None:
        lib.fsph_default_params(C.byref(p), dim)
    else:
        p.dim = dim
        p.sph_kernel, p.eos_type, p.sph_viscosity = 3, 0, 0
        p.sph_eta, p.sph_sinc_index, p.poly_gamma = 1.5, 4.0, 1.4
        p.sph_viscosity_alpha, p.sph_viscosity_beta, p.sph_viscosity_epsilon = 1.0, 2.0, 0.01
        p.timestep_cfl_factor, p.initial_dt, p.gravitational_constant = 0.25, 0.001, 1.0
        p.box_length, p.box_width, p.box_height = 3.0, 1.0, 1.0
        p.evolve_internal_energy = p.thermokinetic_formulation = p.pressure_updates_number = 1
        p.sph_viscosity_alpha_max, p.sph_viscosity_l, p.sph_viscosity_delta = 1.0, 0.05, 1.0
        p.poly_gamma2, p.ppt_density_thr = 2.5, 5e14
        p.relaxation_beta, p.relaxation_gamma, p.relaxation_repulsion_radius, p.relaxation_repulsion_gamma = 1e-6, 0.0, 0.25, 0.0
        p.external_force_type, p.density_profile = b"none", b"constant"
        p.sphere_radius, p.extforce_wall_powerindex, p.extforce_wall_steepness = 1.0, 5.0, 1e12
        p.gravity_acceleration_constant, p.mesa_rim_width, p.rho_initial, p.pressure_initial = 9.81, 0.25, 1.0, 1.0
        p.orbital_separation, p.mass_neutron_star, p.mass_white_dwarf = 2.5e9, 1.26 * 1.988435e33, 1.10 * 1.988435e33
        p.airfoil_size, p.airfoil_thickness, p.airfoil_camber, p.airfoil_anchor_x = 2.0, 0.05, 0.1, -1.0
    for k, v in kw.items():
        if k == "sph_kernel" and isinstance(v, str):
            v = KERNELS[v.lower()]
        if isinstance(v, str):
            v = v.encode()
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class FsphError(RuntimeError):
    pass


class BodySystem:
    """Host-side mirror of body_system<double,D> (mpisph/bodies_system.h) over the C ABI."""

    def __init__(self, params, backend="cuda", device=0):
        self.lib = load(library_path(backend, params.dim))
        self.params = params
        self.dim = params.dim
        self.h = self.lib.fsph_create(C.byref(params), device)
        if not self.h:
            raise FsphError("fsph_create failed for backend %r (no CUDA device, or dimension mismatch)" % backend)
        self.backend = self.lib.fsph_backend().decode()

    def close(self):
        if self.h:
            self.lib.fsph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FsphError(self.lib.fsph_last_error(self.h).decode())

    # -- several ranks (MPI_Comm_rank / size of the reference, bodies_system.h:47-52) ---------------
    def comm_init_nccl(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.fsph_comm_init_nccl(self.h, buf, rank, nranks))

    def comm_init_local(self, group, rank):
        self._ck(self.lib.fsph_comm_init_local(self.h, group, rank))

    def comm_rank(self):
        return int(self.lib.fsph_comm_rank(self.h))

    def comm_size(self):
        return int(self.lib.fsph_comm_size(self.h))

    def nghosts(self):
        return int(self.lib.fsph_nghosts(self.h))

    # -- storage ---------------------------------------------------------
    def resize(self, n):
        self._ck(self.lib.fsph_resize(self.h, int(n)))

    def nbodies(self):
        return int(self.lib.fsph_nbodies(self.h))

    def _shape(self, col):
        n = self.nbodies()
        return (n, self.dim) if col in VECTOR_COLS else (n,)

    def set(self, col, arr):
        a = np.ascontiguousarray(arr, dtype=COL_DTYPE.get(col, np.float64))
        if a.shape != self._shape(col):
            a = a.reshape(self._shape(col))
        self._ck(self.lib.fsph_set_column(self.h, col, a.ctypes.data))

    def get(self, col):
        a = np.empty(self._shape(col), dtype=COL_DTYPE.get(col, np.float64))
        self._ck(self.lib.fsph_get_column(self.h, col, a.ctypes.data))
        return a

    def set_component(self, col, comp, ptr):
        """Contiguous [n] component from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._ck(self.lib.fsph_set_component(self.h, col, comp, ptr))

    def get_component(self, col, comp, ptr):
        self._ck(self.lib.fsph_get_component(self.h, col, comp, ptr))

    def launch_count(self):
        return int(self.lib.fsph_launch_count())

    def load_bodies(self, fields):
        """read_bodies (bodies_system.h:83) from a dict of column-name -> array."""
        n = len(fields["mass"])
        self.resize(n)
        for name, arr in fields.items():
            self.set(COL_NAMES.index(name), arr)

    # -- initial data on the device (lattice.h generators + the sphere set-ups of app/id_generators) ------------
    def generate_lattice(self, lattice_type, domain_type, bmin, bmax, sep, slab=(0, 1), id_offset=0):
        lo, hi = np.ascontiguousarray(bmin, np.float64), np.ascontiguousarray(bmax, np.float64)
        n = C.c_int64()
        self._ck(self.lib.fsph_generate_lattice(self.h, lattice_type, domain_type, lo.ctypes.data, hi.ctypes.data, float(sep),
                                                 int(slab[0]), int(slab[1]), int(id_offset), C.byref(n)))
        return n.value

    def count_within(self, center, radius):
        c = np.ascontiguousarray(center, np.float64)
        n = C.c_int64()
        self._ck(self.lib.fsph_count_within(self.h, c.ctypes.data, float(radius), C.byref(n)))
        return n.value

    def id_fill(self, mass, h, rho, u, u_blast_add=0.0, blast_radius=0.0, gamma=1.4, dt=0.0, radial_velocity=0.0, rho_eos=None, pressure=None):
        """rho_eos: the density P = rho_eos u (gamma - 1) is computed with (default: rho); pressure: a constant instead."""
        f = np.array([mass, h, rho, u, u_blast_add, blast_radius, gamma, dt, radial_velocity, rho if rho_eos is None else rho_eos,
                      float("nan") if pressure is None else pressure], np.float64)
        self._ck(self.lib.fsph_id_fill(self.h, f.ctypes.data))

    def snapshot(self, names=None):
        names = names or [c for c in COL_NAMES if c != "nneigh"]
        return {nm: self.get(COL_NAMES.index(nm)) for nm in names}

    # -- time ------------------------------------------------------------
    def set_time(self, dt, totaltime=0.0, iteration=0):
        self._ck(self.lib.fsph_set_time(self.h, dt, totaltime, iteration))

    def get_time(self):
        dt, t, it = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.lib.fsph_get_time(self.h, C.byref(dt), C.byref(t), C.byref(it)))
        return dt.value, t.value, it.value

    def set_time_state(self, st, iteration):
        """st = [dt, dt_saved, totaltime, totaltime_next, t_h5data_output, t_screen_output, t_scalar_output]."""
        a = np.ascontiguousarray(st, dtype=np.float64)
        self._ck(self.lib.fsph_set_time_state(self.h, a.ctypes.data, int(iteration)))

    def get_time_state(self):
        a = np.zeros(7)
        it = C.c_int64()
        self._ck(self.lib.fsph_get_time_state(self.h, a.ctypes.data, C.byref(it)))
        return a, it.value

    def tree_root(self):
        com = np.zeros(self.dim)
        m = C.c_double()
        self._ck(self.lib.fsph_tree_root(self.h, com.ctypes.data, C.byref(m)))
        return com, m.value

    def advance_time(self):
        self._ck(self.lib.fsph_advance_time(self.h))

    # -- hot path (names of the reference's body_system members) ----------
    def update_iteration(self):
        self._ck(self.lib.fsph_update_iteration(self.h))

    def reset_ghosts(self):
        self._ck(self.lib.fsph_reset_ghosts(self.h))

    def apply_in_smoothinglength(self, op):
        self._ck(self.lib.fsph_sweep(self.h, op))

    def apply_all(self, op):
        self._ck(self.lib.fsph_apply(self.h, op))

    def get_all(self, op, nout=3):
        out = np.zeros(max(nout, 1))
        self._ck(self.lib.fsph_reduce(self.h, op, out.ctypes.data, nout))
        return out

    def gravitation_fmm(self):
        self._ck(self.lib.fsph_gravitation_fmm(self.h))

    def get_range(self):
        lo, hi = np.zeros(self.dim), np.zeros(self.dim)
        self._ck(self.lib.fsph_get_range(self.h, lo.ctypes.data, hi.ctypes.data))
        return lo, hi

    def get_smoothinglength(self):
        return float(self.lib.fsph_get_smoothinglength(self.h))

    def step(self, initial=False):
        self._ck(self.lib.fsph_step(self.h, 1 if initial else 0))

    def step_streamed(self, outputs, initial=False):
        """fsph_step_streamed: `outputs` = [(column, component, address of a pinned host buffer)]."""
        arr = (StreamOut * len(outputs))()
        for k, (col, comp, ptr) in enumerate(outputs):
            arr[k].col, arr[k].comp, arr[k].host = col, comp, ptr
        self._ck(self.lib.fsph_step_streamed(self.h, 1 if initial else 0, C.cast(arr, C.c_void_p), len(outputs)))

    def step_pipelined(self, inputs, outputs, initial=False):
        """fsph_step_pipelined: `inputs` / `outputs` = [(column, component, address of a pinned host buffer)]; returns when
        the kernels are done -- stream_wait() before reading the output buffers."""
        ain, aout = (StreamOut * max(len(inputs), 1))(), (StreamOut * max(len(outputs), 1))()   # fsph_stream_in has the same layout
        for k, (col, comp, ptr) in enumerate(inputs):
            ain[k].col, ain[k].comp, ain[k].host = col, comp, ptr
        for k, (col, comp, ptr) in enumerate(outputs):
            aout[k].col, aout[k].comp, aout[k].host = col, comp, ptr
        self._ck(self.lib.fsph_step_pipelined(self.h, 1 if initial else 0, C.cast(ain, C.c_void_p), len(inputs),
                                              C.cast(aout, C.c_void_p), len(outputs)))

    def stream_wait(self):
        self._ck(self.lib.fsph_stream_wait(self.h))

    # -- introspection ------------------------------------------------------
    def tree_nodes(self):
        nn = int(self.lib.fsph_tree_num_nodes(self.h))
        d = self.dim
        out = dict(key=np.zeros(nn, np.uint64), mass=np.zeros(nn), com=np.zeros((nn, d)), radius=np.zeros(nn),
                   lap=np.zeros(nn), bmin=np.zeros((nn, d)), bmax=np.zeros((nn, d)), sub_entities=np.zeros(nn, np.int64))
        self._ck(self.lib.fsph_tree_get_nodes(self.h, *[out[k].ctypes.data for k in
                                                          ("key", "mass", "com", "radius", "lap", "bmin", "bmax", "sub_entities")]))
        return out

    def leaf_depth(self):
        a = np.zeros(self.nbodies(), np.int32)
        self._ck(self.lib.fsph_tree_get_leaf_depth(self.h, a.ctypes.data))
        return a

    def neighbor_lists(self):
        n = self.nbodies()
        off = np.zeros(n + 1, np.int64)
        self._ck(self.lib.fsph_neighbor_lists(self.h, off.ctypes.data, None))
        ids = np.zeros(int(off[-1]), np.int64)
        self._ck(self.lib.fsph_neighbor_lists(self.h, off.ctypes.data, ids.ctypes.data))
        return off, ids

    def phase_times(self):
        ms = np.zeros(64)
        names = C.c_char_p()
        k = self.lib.fsph_phase_times(self.h, ms.ctypes.data, 64, C.byref(names))
        nm = names.value.decode().split(";") if names.value else []
        return dict(zip(nm[:k], ms[:k].tolist()))


def nccl_unique_id():
    """128 opaque bytes from rank 0 for fsph_comm_init_nccl (ship them to the other ranks, e.g. with torch.distributed)."""
    lib = load(PRODUCT_LIB)
    buf = C.create_string_buffer(128)
    if lib.fsph_comm_unique_id(buf) != 0:
        raise FsphError(lib.fsph_last_error(None).decode())
    return buf.raw


class LocalRanks:
    """P body systems of ONE process joined by the in-process communicator, one host thread per rank (ctypes releases
    the GIL during the calls).  Every collective call of the C ABI is issued on all ranks at once through `call`."""

    def __init__(self, params, nranks, devices=None):
        from concurrent.futures import ThreadPoolExecutor
        self.lib = load(PRODUCT_LIB)
        self.n = nranks
        self.group = self.lib.fsph_local_group_create(nranks)
        devices = devices or [0] * nranks
        self.ranks = [BodySystem(params, backend="cuda", device=devices[r]) for r in range(nranks)]
        for r, s in enumerate(self.ranks):
            s.comm_init_local(self.group, r)
        self.pool = ThreadPoolExecutor(max_workers=nranks)

    def call(self, fn):
        """fn(rank_index, system) on every rank concurrently; returns the list of results."""
        futs = [self.pool.submit(fn, r, s) for r, s in enumerate(self.ranks)]
        errs = [f.exception() for f in futs]
        real = [e for e in errs if e is not None and "local group:" not in str(e)]
        if real:
            raise real[0]  # the rank that failed first; its peers only report the broken group
        for e in errs:
            if e is not None:
                raise e
        return [f.result() for f in futs]

    def close(self):
        for s in self.ranks:
            s.close()
        self.ranks = []
        if self.group:
            self.lib.fsph_local_group_destroy(self.group)
            self.group = None
        self.pool.shutdown()
