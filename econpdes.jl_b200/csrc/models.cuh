// Registered model functors: the `pde(state, u) -> (; vt...)[, (; saved...)]` bodies of
// the reference examples, restated as stateless per-point device functions.
//
// Each functor cites the reference body it follows.  Operand order is kept exactly
// (the translation unit is built with -fmad=false; `x^2` in Julia is x*x).  The
// parameter vector layout of every model is part of the C ABI (include/econpdes_cuda.h,
// `econpdes_model_info`) and mirrored by econpdes.jl_b200/models.py.
//
// Interface of a functor M:
//   M::N, M::F            number of states / unknowns
//   M::NPARAMS, M::NSAVED parameter count, number of saved scalars per point
//   M::USES_CROSS         false if the body never reads a cross derivative: the
//                         Jacobian is then a star stencil and the corner columns of the
//                         +-1 box are exactly zero (the reference stores them as
//                         structural zeros, coloring.jl:18-35)
//   M::eval<SAVE>(prm, state, u, out, saved)
#pragma once
#include <math.h>

#include "stencil.cuh"

namespace econpdes {

EP_HD double ep_min(double a, double b) { return a < b ? a : b; }   // operands are never NaN where used
EP_HD double ep_max(double a, double b) { return a > b ? a : b; }

// README.md:58-67  rho*v = x + mu(x) v_x
struct ReadmeLinear {
  static constexpr int N = 1, F = 1, NPARAMS = 2, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "readme_linear"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<1>* u, double* out, double*) {
    const double rho = prm[0], kappa = prm[1];
    const double x = st[0];
    const double mu = -kappa * (x - 0.5);
    out[0] = -(x + mu * (mu >= 0 ? u[0].up[0] : u[0].dn[0]) - rho * u[0].v);
  }
};

// examples/asset_pricing/campbell_cochrane.jl:66-84
struct CampbellCochrane {
  static constexpr int N = 1, F = 1, NPARAMS = 6, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "campbell_cochrane"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<1>* u, double* out, double*) {
    const double mu = prm[0], sigma = prm[1], gamma = prm[2], rho = prm[3], kappa_s = prm[4], b = prm[5];
    const double s = st[0];
    const double p = u[0].v, pss = u[0].dd[0];
    const double Sbar = sigma * sqrt(gamma / (kappa_s - b / gamma));
    const double sbar = log(Sbar);
    const double lam = 1 / Sbar * sqrt(1 - 2 * (s - sbar)) - 1;
    const double mus = -kappa_s * (s - sbar);
    const double sigs = lam * sigma;
    const double mus_eff = mus + sigs * (sigma - gamma * (sigma + sigs));
    const double ps = (mus_eff >= 0) ? u[0].up[0] : u[0].dn[0];
    const double sigp = ps / p * sigs;
    const double mup = ps / p * mus + 0.5 * pss / p * (sigs * sigs);
    const double kap = gamma * (sigma + sigs);
    const double r = rho + gamma * mu - (gamma * kappa_s - b) / 2 + b * (sbar - s);
    out[0] = -p * (1 / p + mu + mup + sigp * sigma - r - kap * (sigma + sigp));
  }
};

// examples/asset_pricing/bansal_yaron.jl:65-95   states (mu, v)
struct BansalYaron {
  static constexpr int N = 2, F = 1, NPARAMS = 9, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "bansal_yaron"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<2>* u, double* out, double*) {
    const double mubar = prm[0], vbar = prm[1], kappa_mu = prm[2], nu_mu = prm[3], kappa_v = prm[4],
                 nu_v = prm[5], rho = prm[6], gamma = prm[7], psi = prm[8];
    const double mu = st[0], v = st[1];
    const double p = u[0].v;
    const double muc = mu;
    const double sigc = sqrt(v);
    const double mumu = kappa_mu * (mubar - mu);
    const double sigmu = nu_mu * sqrt(v);
    const double muv = kappa_v * (vbar - v);
    const double sigv = nu_v * sqrt(v);
    const double pmu = (mumu >= 0) ? u[0].up[0] : u[0].dn[0];
    const double pv = (muv >= 0) ? u[0].up[1] : u[0].dn[1];
    const double sigp_Zmu = pmu / p * sigmu;
    const double sigp_Zv = pv / p * sigv;
    const double sigp2 = sigp_Zmu * sigp_Zmu + sigp_Zv * sigp_Zv;
    const double mup = pmu / p * mumu + pv / p * muv + 0.5 * u[0].dd[0] / p * (sigmu * sigmu) +
                       0.5 * u[0].dd[1] / p * (sigv * sigv);
    const double kap_Zc = gamma * sigc;
    const double kap_Zmu = -(1 - gamma * psi) / (psi - 1) * sigp_Zmu;
    const double kap_Zv = -(1 - gamma * psi) / (psi - 1) * sigp_Zv;
    const double r = rho + muc / psi - (1 + 1 / psi) / 2 * gamma * (sigc * sigc) -
                     (gamma * psi - 1) / (2 * (psi - 1)) * sigp2;
    out[0] = -p * (1 / p + muc + mup - r - kap_Zc * sigc - kap_Zmu * sigp_Zmu - kap_Zv * sigp_Zv);
  }
};

// examples/consumption_saving/consumption_saving_diffusion_income.jl:71-106  states (y, a)
// saved = (mua, c)
struct AhlmDiffusion {
  static constexpr int N = 2, F = 1, NPARAMS = 8, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "ahlm_diffusion"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<2>* u, double* out, double* sv) {
    const double kappa_y = prm[0], ybar = prm[1], sigma_y = prm[2], r = prm[3], rho = prm[4], gamma = prm[5],
                 amin = prm[6];
    const double y = st[0], a = st[1];
    const double v = u[0].v, vy_up = u[0].up[0], vy_down = u[0].dn[0], va_up = u[0].up[1], va_down = u[0].dn[1],
                 vyy = u[0].dd[0];
    const double muy = kappa_y * (ybar - y);
    const double cmax = 100.0 * (y + r * ep_max(a, 0.0));
    const double vy = (muy >= 0) ? vy_up : vy_down;
    const double c_up = va_up > 0 ? ep_min(pow(va_up, -1 / gamma), cmax) : cmax;
    const double mua_up = y + r * a - c_up;
    double c, va, mua;
    if (mua_up >= 0.0) {
      va = va_up; c = c_up; mua = mua_up;
    } else {
      const double c_down = va_down > 0 ? ep_min(pow(va_down, -1 / gamma), cmax) : cmax;
      const double mua_down = y + r * a - c_down;
      if ((mua_down <= 0.0) && (a > amin)) {
        va = va_down; c = c_down; mua = mua_down;
      } else {
        c = y + r * a;
        va = pow(c, -gamma);
        mua = 0.0;
      }
    }
    out[0] = -(pow(c, 1 - gamma) / (1 - gamma) + mua * va + muy * vy + 0.5 * vyy * (sigma_y * sigma_y) - rho * v);
    if (SAVE) { sv[0] = mua; sv[1] = c; }
  }
};

// Synthetic 3-state HJB of config C5; definition shared verbatim with
// oracle/models.py:hjb3d_synth.
struct Hjb3dSynth {
  static constexpr int N = 3, F = 1, NPARAMS = 9, NSAVED = 0;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "hjb3d_synth"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<3>* u, double* out, double*) {
    const double kx = prm[0], ky = prm[1], kz = prm[2], sx = prm[3], sy = prm[4], sz = prm[5], corr = prm[6],
                 rho = prm[7], theta = prm[8];
    const double x = st[0], y = st[1], z = st[2];
    const double v = u[0].v;
    const double mx = -kx * (x - 0.5), my = -ky * (y - 0.5), mz = -kz * (z - 0.5);
    const double vx = (mx >= 0) ? u[0].up[0] : u[0].dn[0];
    const double vy = (my >= 0) ? u[0].up[1] : u[0].dn[1];
    const double vz = (mz >= 0) ? u[0].up[2] : u[0].dn[2];
    const double drift = mx * vx + my * vy + mz * vz;
    const double diff = 0.5 * (sx * sx) * u[0].dd[0] + 0.5 * (sy * sy) * u[0].dd[1] + 0.5 * (sz * sz) * u[0].dd[2];
    const double cross = corr * sx * sy * u[0].xu[0];
    out[0] = -(x + y + z + drift + diff + cross - rho * v - theta * (v * v));
  }
};

// examples/neoclassical_growth.jl:61-84   saved = (c, muk)
struct NeoclassicalGrowth {
  static constexpr int N = 1, F = 1, NPARAMS = 5, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "neoclassical_growth"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<1>* u, double* out, double* sv) {
    const double A = prm[0], alpha = prm[1], delta = prm[2], rho = prm[3], gamma = prm[4];
    const double k = st[0];
    const double v = u[0].v, vk_up = u[0].up[0], vk_down = u[0].dn[0];
    const double cap = 10 * A * pow(k, alpha);
    const double c_up = vk_up >= 0 ? ep_min(pow(vk_up, -1 / gamma), cap) : cap;
    const double muk_up = A * pow(k, alpha) - delta * k - c_up;
    double c, vk, muk;
    if (muk_up > 0) {
      c = c_up; vk = vk_up; muk = muk_up;
    } else {
      const double c_down = vk_down >= 0 ? ep_min(pow(vk_down, -1 / gamma), cap) : cap;
      const double muk_down = A * pow(k, alpha) - delta * k - c_down;
      if (muk_down < 0) {
        c = c_down; vk = vk_down; muk = muk_down;
      } else {
        muk = 0.0;
        c = A * pow(k, alpha) - delta * k;
        vk = pow(c, -gamma);
      }
    }
    out[0] = -(pow(c, 1 - gamma) / (1 - gamma) + muk * vk - rho * v);
    if (SAVE) { sv[0] = c; sv[1] = muk; }
  }
};

// test/runtests.jl:28-44 (branch structure differs from the example above)
struct GrowthTest {
  static constexpr int N = 1, F = 1, NPARAMS = 5, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "growth_test"; }
  template <bool SAVE>
  EP_HD static void eval(const double* prm, const double* st, const Deriv<1>* u, double* out, double* sv) {
    const double A = prm[0], alpha = prm[1], delta = prm[2], rho = prm[3], gamma = prm[4];
    const double k = st[0];
    const double vk_up = u[0].up[0], vk_down = u[0].dn[0];
    const double cmax = 10 * A * pow(k, alpha);
    const double c_up = vk_up > 0 ? ep_min(pow(vk_up, -1 / gamma), cmax) : cmax;
    const double mu_up = A * pow(k, alpha) - delta * k - c_up;
    const double c_down = vk_down > 0 ? ep_min(pow(vk_down, -1 / gamma), cmax) : cmax;
    const double mu_down = A * pow(k, alpha) - delta * k - c_down;
    double c, vk, muk;
    if (mu_up > 0) {
      c = c_up; vk = vk_up; muk = mu_up;
    } else if (mu_down < 0) {
      c = c_down; vk = vk_down; muk = mu_down;
    } else {
      c = A * pow(k, alpha) - delta * k;
      vk = pow(c, -gamma);
      muk = 0.0;
    }
    out[0] = -(pow(c, 1 - gamma) / (1 - gamma) + muk * vk - rho * u[0].v);
    if (SAVE) { sv[0] = c; sv[1] = muk; }
  }
};

// test/runtests.jl:276-279, 314-317:  vt = -(v - x - z - ...)
template <int NDIM>
struct LinearSum {
  static constexpr int N = NDIM, F = 1, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return NDIM == 1 ? "linear_sum_1d" : (NDIM == 2 ? "linear_sum_2d" : "linear_sum_3d"); }
  template <bool SAVE>
  EP_HD static void eval(const double*, const double* st, const Deriv<NDIM>* u, double* out, double*) {
    double acc = u[0].v;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) acc = acc - st[d];
    out[0] = -acc;
  }
};

// test/runtests.jl:379-384: two unknowns (v, p): v = x + p, p = v/2 written forward
struct ForwardPrice {
  static constexpr int N = 1, F = 2, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "forward_price"; }
  template <bool SAVE>
  EP_HD static void eval(const double*, const double* st, const Deriv<1>* u, double* out, double*) {
    out[0] = -(st[0] + u[1].v - u[0].v);
    out[1] = u[0].v / 2 - u[1].v;
  }
};

// test/runtests.jl:227: vt = v - 2 (used with lower_bound for the MCP known answer)
struct VMinusTwo {
  static constexpr int N = 1, F = 1, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "v_minus_two"; }
  template <bool SAVE>
  EP_HD static void eval(const double*, const double*, const Deriv<1>* u, double* out, double*) {
    out[0] = u[0].v - 2.0;
  }
};

// X-macro: the registry order defines the model ids of the C ABI.
#define ECONPDES_FOR_EACH_MODEL(X) \
  X(ReadmeLinear)                  \
  X(CampbellCochrane)              \
  X(BansalYaron)                   \
  X(AhlmDiffusion)                 \
  X(Hjb3dSynth)                    \
  X(NeoclassicalGrowth)            \
  X(GrowthTest)                    \
  X(LinearSum<1>)                  \
  X(LinearSum<2>)                  \
  X(LinearSum<3>)                  \
  X(ForwardPrice)                  \
  X(VMinusTwo)

}  // namespace econpdes
