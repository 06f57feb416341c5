"""Inputs shared by the external-force tests.  One value per parameter for the whole test session: the reference's terms
keep them in function-level statics after their first call (eforce.h:51-55,73-77,141-142,191,294-296,261-263)."""
import numpy as np

EF_PARAMS = dict(
    box_length=1.0, box_width=0.8, box_height=0.6, extforce_wall_powerindex=3.0, extforce_wall_steepness=40.0,
    sphere_radius=0.55, gravity_acceleration_constant=0.7, zero_potential_poison_value=0.125, mesa_rim_width=0.3,
    rho_initial=1.3, pressure_initial=0.9, poly_gamma=1.5, gravitational_constant=1.0, orbital_separation=2.0,
    mass_neutron_star=0.6, mass_white_dwarf=0.4, airfoil_size=0.8, airfoil_thickness=0.2, airfoil_camber=0.05,
    airfoil_anchor_x=-0.3, airfoil_anchor_y=0.02, airfoil_attack_angle=7.0)

# name -> (dimension, external_force_type, extra parameters).  density_profile is not latched (density_profiles::select
# re-points the function pointers), but K0 = pressure_initial / rho_initial^gamma and rho0 = profile(0) are, at the first
# call of the density-support term: one profile per dimension.
CASES = {
    "walls_gravity_3d": (3, "walls:xyz,gravity", {}),
    "sphere_orbit_poison_3d": (3, "spherical wall,orbit,poison", {}),
    "support_parabolic_3d": (3, "spherical density support", dict(density_profile="parabolic")),
    "airfoil_walls_2d": (2, "airfoil,walls:xy", {}),
    "airfoil_2d": (2, "airfoil", {}),
    "support_mesa_gravity_2d": (2, "spherical density support,gravity", dict(density_profile="mesa")),
    "walls_gravity_1d": (1, "walls:x,gravity", {}),
}


def points(dim, n, seed):
    """Points in [-0.7, 0.7]^dim: some beyond every wall of EF_PARAMS, some inside the airfoil."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.7, 0.7, (n, dim))
    x[0] = 0.05   # near the centre
    return x
