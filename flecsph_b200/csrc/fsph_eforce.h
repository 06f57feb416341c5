// fsph_eforce.h -- external forces and potentials (include/physics/eforce.h) and the spherical density profiles the
// "spherical density support" term reads (include/physics/density_profiles.h:69-215,355-396).
//
// external_force::select (eforce.h:358-433) turns the comma-separated `external_force_type` keyword into two lists of
// function pointers; external_force::acceleration / potential (:330-352) add the terms up in list order, starting from
// zero.  Here the keyword is parsed once on the host (ef_parse) into a list of term codes that travels to the kernels by
// value; ef_acceleration / ef_potential evaluate the terms in the same order with the reference's expressions.  Used by
// the element-wise pass and the energy reduction (fsph_apply.cu), which are compiled without FMA contraction.
#pragma once
#include <cmath>
#include <cctype>
#include <cstring>
#include <string>
#include <stdexcept>
#include "flecsph_b200.h"

namespace fsph {

enum EfTerm {
  EF_SPHERICAL_WALL = 0,   // eforce.h:92-131
  EF_AIRFOIL,              // :222-282
  EF_DENSITY_SUPPORT,      // :138-180 (includes the spherical wall)
  EF_GRAVITY,              // :188-205
  EF_ORBIT,                // :290-330
  EF_WALL_X, EF_WALL_Y, EF_WALL_Z,   // :46-86
  EF_POISON                // :336-339 (potential only)
};
enum EfProfile { EF_PROFILE_CONSTANT = 0, EF_PROFILE_PARABOLIC, EF_PROFILE_MESA };

struct ExtForce {
  int n = 0;         // terms, in the order of the keyword
  int nacc = 0;      // terms that contribute an acceleration (all but "poison")
  int term[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int profile = EF_PROFILE_CONSTANT;
  double box[3] = {0, 0, 0};             // half extents: .5 * box_length / box_width / box_height
  double pw_n = 5.0, pw_a = 1e12;        // extforce_wall_powerindex / _steepness
  double sphere_radius = 1.0, grav = 9.81, poison = 0.0;
  double gamma = 1.4, rho_initial = 1.0, K0 = 1.0, rho0 = 1.0, mesa_q = 0.25, mesa_rho0 = 0.0;
  double gc = 1.0, a_sp = 0.0, m_ns = 0.0, m_wd = 0.0;
  double af_size = 2.0, af_thick = 0.05, af_camber = 0.1, af_x = -1.0, af_y = 0.0, af_alpha = 0.0;
};

#define EF_SQ(x) ((x) * (x))
#define EF_CU(x) ((x) * (x) * (x))

// density_profiles::mesa_mass_helper, density_profiles.h:137-151
template<int D>
__host__ __device__ inline double ef_mesa_mass_helper(double q, double r) {
  const double dr = q, r0 = 1. - q;
  double mm = 0.0;
  if(D == 2) mm = EF_SQ(r) * (1. - (.5 * EF_SQ(r) + EF_SQ(r0)) / EF_SQ(dr)) + (.5 * EF_CU(r0) + 4. * EF_CU(r)) * r0 / (3. * EF_SQ(dr));
  if(D == 3)
    mm = EF_CU(r0) / 3. + (EF_CU(r) - EF_CU(r0)) / 3. * (1. - EF_SQ(r0) / EF_SQ(dr)) +
         r0 * (EF_SQ(r) - EF_SQ(r0)) * (EF_SQ(r) + EF_SQ(r0)) / (2. * EF_SQ(dr)) - (EF_SQ(r) * EF_CU(r) - EF_SQ(r0) * EF_CU(r0)) / (5 * EF_SQ(dr));
  return mm;
}

// spherical_density_profile(r): rho_constant_density :69, rho_parabolic_density :107, rho_mesa_density :179-188
template<int D>
__host__ __device__ inline double ef_profile_rho(const ExtForce & ef, double r) {
  const double pi = 3.14159265358979323846;
  if(ef.profile == EF_PROFILE_PARABOLIC) return D == 1 ? 0.75 * (1. - EF_SQ(r)) : (D == 2 ? 2. / pi * (1. - EF_SQ(r)) : 15. / (8. * pi) * (1. - EF_SQ(r)));
  if(ef.profile == EF_PROFILE_MESA) {
    const double r0 = 1. - ef.mesa_q;
    if(r < 1. - ef.mesa_q) return ef.mesa_rho0;
    if(r < .9999) return ef.mesa_rho0 * (1. - EF_SQ(r - r0) / EF_SQ(ef.mesa_q));
    return 0.0;
  }
  return D == 1 ? 0.5 : (D == 2 ? 1.0 / pi : 0.75 / pi);
}

// spherical_drho_dr(r): drhodr_constant_density :98, drhodr_parabolic_density :122, drhodr_mesa_density :214-221
template<int D>
__host__ __device__ inline double ef_profile_drhodr(const ExtForce & ef, double r) {
  const double pi = 3.14159265358979323846;
  if(ef.profile == EF_PROFILE_PARABOLIC) return D == 1 ? -1.5 * r : (D == 2 ? -4. / pi * r : -15. / (4. * pi) * r);
  if(ef.profile == EF_PROFILE_MESA) {
    const double r0 = 1. - ef.mesa_q;
    if(r > 1. - ef.mesa_q && r < .9999) return -2. * ef.mesa_rho0 * (r - r0) / EF_SQ(ef.mesa_q);
    return 0.0;
  }
  return 0.0;
}

// component I of a D-vector without a dynamically indexed load (keeps small arrays in registers)
template<int D>
__host__ __device__ inline double ef_pick(const double * x, int I) {
  return I == 0 ? x[0] : (I == 1 ? x[D > 1 ? 1 : 0] : x[D > 2 ? 2 : 0]);
}

template<int D>
__host__ __device__ inline double ef_radius(const double * x) {
  double r = x[0] * x[0];
#pragma unroll
  for(int i = 1; i < D; ++i) r += x[i] * x[i];
  return sqrt(r);
}

// acceleration_spherical_wall, eforce.h:92-109
template<int D>
__host__ __device__ inline void ef_acc_spherical_wall(const ExtForce & ef, const double * x, double * a) {
  const double r = ef_radius<D>(x);
  if(r > ef.sphere_radius) {
    const double ar = ef.pw_n * ef.pw_a * pow(r - ef.sphere_radius, ef.pw_n - 1);
#pragma unroll
    for(int i = 0; i < D; ++i) a[i] += -x[i] / r * ar;
  }
}

// potential_spherical_wall, eforce.h:111-125
template<int D>
__host__ __device__ inline double ef_pot_spherical_wall(const ExtForce & ef, const double * x) {
  const double r = ef_radius<D>(x);
  return r > ef.sphere_radius ? ef.pw_a * pow(r - ef.sphere_radius, ef.pw_n) : 0.0;
}

// external_force::acceleration (eforce.h:330-337): a = 0, then a += term(particle) for every term of the keyword.
// A term's own vector is built first (components it does not touch are 0) and then added, as in the reference.
template<int D>
__host__ __device__ inline void ef_acceleration(const ExtForce & ef, const double * x, double * a) {
#pragma unroll
  for(int i = 0; i < D; ++i) a[i] = 0.0;
  for(int k = 0; k < ef.n; ++k) {
    double t[3] = {0.0, 0.0, 0.0};
    const int term = ef.term[k];
    switch(term) {
      case EF_SPHERICAL_WALL: ef_acc_spherical_wall<D>(ef, x, t); break;
      case EF_AIRFOIL:
        if(D > 1) {
          // acceleration_airfoil, eforce.h:222-255
          const double pi = 3.14159265358979323846;
          const double x1 = x[0] - ef.af_x, y1 = x[D > 1 ? 1 : 0] - ef.af_y, alpha = ef.af_alpha * pi / 180.0;
          const double ca = cos(alpha), sa = sin(alpha);
          const double xx = x1 * ca + y1 * sa, yy = -x1 * sa + y1 * ca;
          const bool inside = fabs(yy) < 5.0 * ef.af_thick && xx > -ef.af_size * 0.02 && xx < ef.af_size * 1.02;
          const double upper = ef.af_thick * xx * sqrt(ef.af_size * ef.af_size - xx * xx);
          const double camber = ef.af_camber * sin(pi * xx / 2.);
          const double phi = EF_SQ(upper) - EF_SQ(yy - camber) + 0.002;
          if(inside && phi > 0.0) {
            const double a0 = ef.pw_n * ef.pw_a * pow(phi, ef.pw_n - 1) *
                              (2. * (yy - camber) * (-ef.af_camber * pi / 2. * cos(pi / 2. * xx)) -
                                ef.af_thick * ef.af_thick * 2 * xx * (ef.af_size * ef.af_size - 2 * xx * xx));
            const double a1 = ef.pw_n * ef.pw_a * pow(phi, ef.pw_n - 1) * 2. * (yy - camber);
            t[0] = a0 * ca - a1 * sa;
            t[1] = a0 * sa + a1 * ca;
          }
        }
        break;
      case EF_DENSITY_SUPPORT: {
        // acceleration_spherical_density_support, eforce.h:138-160: pressure-gradient support of the initial profile
        // plus the spherical wall
        const double r = ef_radius<D>(x);
        const double q = r / ef.sphere_radius;
        if(q > 1e-12) {
          const double rho = ef.rho_initial / ef.rho0 * ef_profile_rho<D>(ef, q);
          const double drhodr = ef.rho_initial / (ef.rho0 * ef.sphere_radius) * ef_profile_drhodr<D>(ef, q);
          const double a_r = rho > 0 ? ef.K0 * ef.gamma * pow(rho, ef.gamma - 2) * drhodr : 0;
#pragma unroll
          for(int i = 0; i < D; ++i) t[i] = a_r * x[i] / r;
        }
        double w[3] = {0.0, 0.0, 0.0};
        ef_acc_spherical_wall<D>(ef, x, w);
#pragma unroll
        for(int i = 0; i < D; ++i) t[i] = t[i] + w[i];
      } break;
      case EF_GRAVITY: t[D > 1 ? 1 : 0] = -ef.grav; break;   // eforce.h:188-196
      case EF_ORBIT:
        if(D == 3) {
          // acceleration_orbit, eforce.h:290-313
          const double m_t = ef.m_ns + ef.m_wd;
          double temp = EF_SQ(x[0] - ef.a_sp) + EF_SQ(x[D > 1 ? 1 : 0]) + EF_SQ(x[D > 2 ? 2 : 0]);
          temp = EF_CU(temp);
          temp = sqrt(temp);
          const double term1 = -ef.gc * ef.m_ns / temp;
          t[0] += term1 * (x[0] - ef.a_sp);
          t[1] += term1 * x[D > 1 ? 1 : 0];
          t[2] += term1 * x[D > 2 ? 2 : 0];
          const double term2 = ef.gc * m_t / EF_CU(ef.a_sp);
          t[0] += term2 * (x[0] - ef.a_sp * ef.m_ns / m_t);
          t[1] += term2 * x[D > 1 ? 1 : 0];
        }
        break;
      case EF_WALL_X:
      case EF_WALL_Y:
      case EF_WALL_Z: {
        // acceleration_square_well<I>, eforce.h:66-80
        const int I = term - EF_WALL_X;
        if(I < D) {
          const double xi = ef_pick<D>(x, I), b = ef_pick<3>(ef.box, I);
          const double w = ((xi < -b ? pow(-xi - b, ef.pw_n - 1) : 0.0) - (xi > b ? pow(xi - b, ef.pw_n - 1) : 0.0)) * ef.pw_n * ef.pw_a;
          if(I == 0) t[0] = w;
          else if(I == 1) t[1] = w;
          else t[2] = w;
        }
      } break;
      default: continue;   // poison: no acceleration
    }
#pragma unroll
    for(int i = 0; i < D; ++i) a[i] += t[i];
  }
}

// external_force::potential (eforce.h:343-352)
template<int D>
__host__ __device__ inline double ef_potential(const ExtForce & ef, const double * x) {
  double phi = 0.0;
  for(int k = 0; k < ef.n; ++k) {
    const int term = ef.term[k];
    double p = 0.0;
    switch(term) {
      case EF_SPHERICAL_WALL: p = ef_pot_spherical_wall<D>(ef, x); break;
      case EF_AIRFOIL:
        if(D > 1) {
          // potential_airfoil, eforce.h:257-282
          const double pi = 3.14159265358979323846;
          const double alpha = ef.af_alpha * pi / 180.0;
          const double x1 = x[0] - ef.af_x, y1 = x[D > 1 ? 1 : 0] - ef.af_y;
          const double xx = x1 * cos(alpha) + y1 * sin(alpha), yy = -x1 * sin(alpha) + y1 * cos(alpha);
          const bool inside = fabs(yy) < 5.0 * ef.af_thick && xx > -ef.af_size * 0.02 && xx < ef.af_size * 1.02;
          const double upper = ef.af_thick * xx * sqrt(ef.af_size * ef.af_size - xx * xx);
          const double camber = ef.af_camber * sin(pi * xx / 2.);
          const double aux = EF_SQ(upper) - EF_SQ(yy - camber) + 0.002;
          if(inside && aux > 0.0) p = ef.pw_a * pow(aux, ef.pw_n);
        }
        break;
      case EF_DENSITY_SUPPORT: {
        // potential_spherical_density_support, eforce.h:162-180
        const double r = ef_radius<D>(x);
        const double q = r / ef.sphere_radius;
        const double rho = ef.rho_initial / ef.rho0 * ef_profile_rho<D>(ef, q);
        p = rho > 0 ? -ef.K0 * ef.gamma * pow(rho, ef.gamma - 1.) / (ef.gamma - 1.) : 0;
        p = p + ef_pot_spherical_wall<D>(ef, x);
      } break;
      case EF_GRAVITY: p = x[D > 1 ? 1 : 0] * ef.grav; break;   // eforce.h:198-205
      case EF_ORBIT:
        if(D == 3) {
          // potential_orbit, eforce.h:315-330
          const double m_t = ef.m_ns + ef.m_wd;
          double term1 = sqrt(EF_SQ(x[0] - ef.a_sp) + EF_SQ(x[D > 1 ? 1 : 0]) + EF_SQ(x[D > 2 ? 2 : 0]));
          term1 = -ef.gc * ef.m_ns / term1;
          double term2 = -0.5 * ef.gc * m_t / EF_CU(ef.a_sp);
          term2 = term2 * (EF_SQ(x[0] - ef.a_sp * ef.m_ns / m_t) + EF_SQ(x[D > 1 ? 1 : 0]));
          p = term1 + term2;
        }
        break;
      case EF_WALL_X:
      case EF_WALL_Y:
      case EF_WALL_Z: {
        // potential_square_well<I>, eforce.h:46-60
        const int I = term - EF_WALL_X;
        if(I < D) {
          const double xi = ef_pick<D>(x, I), b = ef_pick<3>(ef.box, I);
          p = ((xi < -b ? pow(-xi - b, ef.pw_n) : 0.0) + (xi > b ? pow(xi - b, ef.pw_n) : 0.0)) * ef.pw_a;
        }
      } break;
      case EF_POISON: p = ef.poison; break;
    }
    phi += p;
  }
  return phi;
}

// ---- host: external_force::select (eforce.h:358-433) + density_profiles::select (density_profiles.h:355-396) ------
inline bool ef_ieq(const std::string & a, const char * b) {
  size_t nb = 0;
  while(b[nb]) ++nb;
  if(a.size() != nb) return false;
  for(size_t i = 0; i < nb; ++i)
    if(std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}

inline ExtForce ef_parse(const fsph_params & p) {
  ExtForce ef;
  const int D = p.dim;
  std::string s(p.external_force_type, strnlen(p.external_force_type, sizeof(p.external_force_type)));
  ef.box[0] = .5 * p.box_length;
  ef.box[1] = .5 * p.box_width;
  ef.box[2] = .5 * p.box_height;
  ef.pw_n = p.extforce_wall_powerindex;
  ef.pw_a = p.extforce_wall_steepness;
  ef.sphere_radius = p.sphere_radius;
  ef.grav = p.gravity_acceleration_constant;
  ef.poison = p.zero_potential_poison_value;
  ef.gamma = p.poly_gamma;
  ef.rho_initial = p.rho_initial;
  ef.gc = p.gravitational_constant;
  ef.a_sp = p.orbital_separation;
  ef.m_ns = p.mass_neutron_star;
  ef.m_wd = p.mass_white_dwarf;
  ef.af_size = p.airfoil_size;
  ef.af_thick = p.airfoil_thickness;
  ef.af_camber = p.airfoil_camber;
  ef.af_x = p.airfoil_anchor_x;
  ef.af_y = p.airfoil_anchor_y;
  ef.af_alpha = p.airfoil_attack_angle;
  if(s.empty() || ef_ieq(s, "zero") || ef_ieq(s, "none")) return ef;
  auto push = [&](int t) {
    if(ef.n >= 8) throw std::runtime_error("external_force_type: more than 8 terms");
    ef.term[ef.n++] = t;
    if(t != EF_POISON) ef.nacc++;
  };
  size_t pos = 0;
  while(true) {
    const size_t c = s.find(',', pos);
    const std::string it = s.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
    if(ef_ieq(it, "spherical wall")) push(EF_SPHERICAL_WALL);
    else if(ef_ieq(it, "airfoil")) {
      if(D < 2) throw std::runtime_error("external_force_type: airfoil needs 2 or 3 dimensions (eforce.h:225)");
      push(EF_AIRFOIL);
    }
    else if(ef_ieq(it, "spherical density support")) {
      std::string prof(p.density_profile, strnlen(p.density_profile, sizeof(p.density_profile)));
      if(prof.empty() || ef_ieq(prof, "constant")) ef.profile = EF_PROFILE_CONSTANT;
      else if(ef_ieq(prof, "parabolic")) ef.profile = EF_PROFILE_PARABOLIC;
      else if(ef_ieq(prof, "mesa")) {
        ef.profile = EF_PROFILE_MESA;
        ef.mesa_q = p.mesa_rim_width;
        const double pi = 3.14159265358979323846;
        if(D == 1) ef.mesa_rho0 = .5 / (1. - ef.mesa_q / 3.);
        if(D == 2) ef.mesa_rho0 = 1. / (pi * ef_mesa_mass_helper<2>(ef.mesa_q, 1.));
        if(D == 3) ef.mesa_rho0 = 1. / (4. * pi * ef_mesa_mass_helper<3>(ef.mesa_q, 1.));
      }
      else throw std::runtime_error("density_profile '" + prof + "' is not available on the device (constant, parabolic, mesa)");
      ef.K0 = p.pressure_initial / pow(p.rho_initial, p.poly_gamma);
      ef.rho0 = D == 1 ? ef_profile_rho<1>(ef, 0.) : (D == 2 ? ef_profile_rho<2>(ef, 0.) : ef_profile_rho<3>(ef, 0.));
      push(EF_DENSITY_SUPPORT);
    }
    else if(ef_ieq(it, "gravity")) push(EF_GRAVITY);
    else if(ef_ieq(it, "orbit")) {
      if(D < 3) throw std::runtime_error("external_force_type: orbit needs 3 dimensions (eforce.h:302-304)");
      push(EF_ORBIT);
    }
    else if(it.size() >= 6 && ef_ieq(it.substr(0, 6), "walls:")) {
      const std::string ax = it.substr(6);
      for(size_t i = 0; i < ax.size() && i < 3; ++i) {
        const int ch = std::tolower((unsigned char)ax[i]);
        if(ch < 'x' || ch > 'z') throw std::runtime_error("bad external_force_type: " + it);
        if(ch - 'x' >= D) throw std::runtime_error("external_force_type: " + it + " names an axis beyond the dimension");
        push(EF_WALL_X + (ch - 'x'));
      }
    }
    else if(ef_ieq(it, "poison")) push(EF_POISON);
    else throw std::runtime_error("bad external_force_type: '" + it + "'");
    if(c == std::string::npos) break;
    pos = c + 1;
  }
  return ef;
}

}  // namespace fsph
