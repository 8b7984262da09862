// Registered model functors: the `pde(state, u) -> (; vt...)[, (; saved...)]` bodies of
// the reference examples, restated as stateless per-point device functions.
//
// Each functor cites the reference body it follows.  Operand order is kept exactly
// (the translation unit is built with -fmad=false; `x^2` in Julia is x*x).  The
// parameter vector layout of every model is part of the C ABI (include/econpdes_cuda.h,
// `econpdes_model_info`) and mirrored by econpdes.jl_b200/models.py.
//
// A functor M is split in two so that kernels which evaluate the model several times at
// one grid point (the colored-FD Jacobian: one evaluation per stencil column; the fused
// batched solve: every Newton iteration) pay the state-only arithmetic once:
//   M::prepare(prm, state, pre)   everything that depends on the grid point and the
//                                 parameters but not on the unknowns (drifts, volatilities,
//                                 sqrt/log/pow of the state ...)
//   M::eval<SAVE>(pre, u, out, saved)   the rest
// The split only moves sub-expressions; every floating-point operation and its operand
// order is the reference's, so the result is bit-identical to evaluating the body in one go.
//
//   M::N, M::F            number of states / unknowns
//   M::NPARAMS, M::NSAVED parameter count, number of saved scalars per point
//   M::USES_CROSS         false if the body never reads a cross derivative: the
//                         Jacobian is then a star stencil and the corner columns of the
//                         +-1 box are exactly zero (the reference stores them as
//                         structural zeros, coloring.jl:18-35)
#pragma once
#include <math.h>

#include "stencil.cuh"

namespace econpdes {

// Defaults shared by the functors.  slot_used(k) says whether the residual can depend on box
// slot k = (o1+1) + 3(o2+1) + 9(o3+1) at all; the Jacobian kernel skips the perturbed evaluation
// for slots a model never reads (their entries are exact zeros in the reference too, because
// (f(x1)-f(x))/eps = 0/eps).
struct ModelDefaults {
  // A model whose body divides by its unknown may offer eval_tame: the same body with the select-free quotients
  // (ep_rcp_tame, mdiv_tame), valid when the unknown is positive and of moderate magnitude (2^-200 <= y < 2^199) and
  // the stencil operands are tame (sweep.cuh).  Same bits as eval there; the fused batched kernel takes it then.
  static constexpr bool HAS_TAME_EVAL = false;
  EP_HD static constexpr bool slot_used(int) { return true; }
  // Parameter-only sub-expressions (quotients of parameters, reciprocals for mdiv) are evaluated once
  // per problem -- on the host when the problem is created, by one thread in the batched kernel -- and
  // appended to the parameter vector at prm[NPARAMS .. NPARAMS + NDERIVED): the same IEEE operations
  // the reference performs at every grid point, so the bits are the same.
  static constexpr int NDERIVED = 0;
  EP_HD static void derive(double*) {}
  // Bodies of the three-argument form pde(state, u, t) (src/pdesolve.jl:253-260) read the time from
  // prm[TIME_SLOT]; the backward-induction driver stores the current time there before every step.
  // -1: the body takes no time argument.
  static constexpr int TIME_SLOT = -1;
};
constexpr int kMaxParams = 24;   // ParamPack: NPARAMS + NDERIVED (+ the time slot)
constexpr int kMaxBatchParams = 16;   // the fused batched kernel stages that many in shared memory

EP_HD double ep_min(double a, double b) { return a < b ? a : b; }   // operands are never NaN where used
EP_HD double ep_max(double a, double b) { return a > b ? a : b; }

// README.md:58-67  rho*v = x + mu(x) v_x
struct ReadmeLinear : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 2, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "readme_linear"; }
  struct Pre { double x, mu, rho; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double rho = prm[0], kappa = prm[1];
    c.x = st[0];
    c.mu = -kappa * (c.x - 0.5);
    c.rho = rho;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    out[0] = -(c.x + c.mu * (c.mu >= 0 ? u[0].up[0] : u[0].dn[0]) - c.rho * u[0].v);
  }
};

// examples/asset_pricing/campbell_cochrane.jl:66-84
struct CampbellCochrane : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 6, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "campbell_cochrane"; }
  struct Pre { double mu, sigma, mus, sigs, sigs2, mus_eff, kap, r; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double mu = prm[0], sigma = prm[1], gamma = prm[2], rho = prm[3], kappa_s = prm[4], b = prm[5];
    const double s = st[0];
    const double Sbar = sigma * sqrt(gamma / (kappa_s - b / gamma));
    const double sbar = log(Sbar);
    const double lam = 1 / Sbar * sqrt(1 - 2 * (s - sbar)) - 1;
    c.mu = mu;
    c.sigma = sigma;
    c.mus = -kappa_s * (s - sbar);
    c.sigs = lam * sigma;
    c.sigs2 = c.sigs * c.sigs;
    c.mus_eff = c.mus + c.sigs * (sigma - gamma * (sigma + c.sigs));
    c.kap = gamma * (sigma + c.sigs);
    c.r = rho + gamma * mu - (gamma * kappa_s - b) / 2 + b * (sbar - s);
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double p = u[0].v, pss = u[0].dd[0];
    const double ps = (c.mus_eff >= 0) ? u[0].up[0] : u[0].dn[0];
    // one IEEE reciprocal of p; the other quotients by p reuse it (mdiv: same bits as `/`)
    const double rp = ep_rcp(p);   // RN(1/p), branch-free
    const double ps_p = mdiv(ps, p, rp);
    const double sigp = ps_p * c.sigs;
    const double mup = ps_p * c.mus + mdiv(0.5 * pss, p, rp) * c.sigs2;
    out[0] = -p * (rp + c.mu + mup + sigp * c.sigma - c.r - c.kap * (c.sigma + sigp));
  }
  static constexpr bool HAS_TAME_EVAL = true;
  template <bool SAVE>
  EP_HD static void eval_tame(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double p = u[0].v, pss = u[0].dd[0];
    const double ps = (c.mus_eff >= 0) ? u[0].up[0] : u[0].dn[0];
    const double rp = ep_rcp_tame(p);
    const double ps_p = mdiv_tame(ps, p, rp);
    const double sigp = ps_p * c.sigs;
    const double mup = ps_p * c.mus + mdiv_tame(0.5 * pss, p, rp) * c.sigs2;
    out[0] = -p * (rp + c.mu + mup + sigp * c.sigma - c.r - c.kap * (c.sigma + sigp));
  }
};

// examples/asset_pricing/bansal_yaron.jl:65-95   states (mu, v)
struct BansalYaron : ModelDefaults {
  static constexpr int N = 2, F = 1, NPARAMS = 9, NSAVED = 0, NDERIVED = 4;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "bansal_yaron"; }
  EP_HD static void derive(double* prm) {
    const double gamma = prm[7], psi = prm[8];
    prm[9] = -(1 - gamma * psi) / (psi - 1);            // cz
    prm[10] = (gamma * psi - 1) / (2 * (psi - 1));      // cs
    prm[11] = 1 / psi;                                  // RN(1/psi) for muc / psi
    prm[12] = (1 + 1 / psi) / 2 * gamma;                // leading factor of the precautionary term
  }
  struct Pre { double muc, mumu, sigmu, sigmu2, muv, sigv, sigv2, cz, r_base, cs, kzc_sigc; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double mubar = prm[0], vbar = prm[1], kappa_mu = prm[2], nu_mu = prm[3], kappa_v = prm[4],
                 nu_v = prm[5], rho = prm[6], gamma = prm[7], psi = prm[8];
    const double mu = st[0], v = st[1];
    c.muc = mu;
    const double sigc = sqrt(v);
    c.mumu = kappa_mu * (mubar - mu);
    c.sigmu = nu_mu * sqrt(v);
    c.sigmu2 = c.sigmu * c.sigmu;
    c.muv = kappa_v * (vbar - v);
    c.sigv = nu_v * sqrt(v);
    c.sigv2 = c.sigv * c.sigv;
    const double kap_Zc = gamma * sigc;
    c.cz = prm[9];
    c.r_base = rho + mdiv(c.muc, psi, prm[11]) - prm[12] * (sigc * sigc);
    c.cs = prm[10];
    c.kzc_sigc = kap_Zc * sigc;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<2>* u, double* out, double*) {
    const double p = u[0].v;
    const double pmu = (c.mumu >= 0) ? u[0].up[0] : u[0].dn[0];
    const double pv = (c.muv >= 0) ? u[0].up[1] : u[0].dn[1];
    const double rp = ep_rcp(p);   // RN(1/p); the other quotients by p reuse it (mdiv: same bits as `/`)
    const double pmu_p = mdiv(pmu, p, rp), pv_p = mdiv(pv, p, rp);
    const double sigp_Zmu = pmu_p * c.sigmu;
    const double sigp_Zv = pv_p * c.sigv;
    const double sigp2 = sigp_Zmu * sigp_Zmu + sigp_Zv * sigp_Zv;
    const double mup = pmu_p * c.mumu + pv_p * c.muv + mdiv(0.5 * u[0].dd[0], p, rp) * c.sigmu2 +
                       mdiv(0.5 * u[0].dd[1], p, rp) * c.sigv2;
    const double kap_Zmu = c.cz * sigp_Zmu;
    const double kap_Zv = c.cz * sigp_Zv;
    const double r = c.r_base - c.cs * sigp2;
    out[0] = -p * (rp + c.muc + mup - r - c.kzc_sigc - kap_Zmu * sigp_Zmu - kap_Zv * sigp_Zv);
  }
};

// examples/consumption_saving/consumption_saving_diffusion_income.jl:71-106  states (y, a)
// saved = (mua, c)
struct AhlmDiffusion : ModelDefaults {
  static constexpr int N = 2, F = 1, NPARAMS = 8, NSAVED = 2, NDERIVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "ahlm_diffusion"; }
  EP_HD static void derive(double* prm) {
    const double gamma = prm[5];
    prm[8] = -1 / gamma;
    prm[9] = 1 / (1 - gamma);   // RN(1/(1 - gamma)) for the quotients by 1 - gamma
  }
  struct Pre { double muy, cmax, inc, va3, util3, neg_inv_gamma, one_m_gamma, r_one_m_gamma, sig2, rho; bool above_amin; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double kappa_y = prm[0], ybar = prm[1], sigma_y = prm[2], r = prm[3], rho = prm[4], gamma = prm[5],
                 amin = prm[6];
    const double y = st[0], a = st[1];
    c.muy = kappa_y * (ybar - y);
    c.cmax = 100.0 * (y + r * ep_max(a, 0.0));
    c.inc = y + r * a;                       // c = y + r a when the drift is pinned at zero
    c.neg_inv_gamma = prm[8];
    c.one_m_gamma = 1 - gamma;
    c.r_one_m_gamma = prm[9];
    c.va3 = pow(c.inc, -gamma);
    c.util3 = mdiv(pow(c.inc, c.one_m_gamma), c.one_m_gamma, c.r_one_m_gamma);
    c.sig2 = sigma_y * sigma_y;
    c.rho = rho;
    c.above_amin = a > amin;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& p, const Deriv<2>* u, double* out, double* sv) {
    const double v = u[0].v, vy_up = u[0].up[0], vy_down = u[0].dn[0], va_up = u[0].up[1], va_down = u[0].dn[1],
                 vyy = u[0].dd[0];
    const double vy = (p.muy >= 0) ? vy_up : vy_down;
    const double c_up = va_up > 0 ? ep_min(pow(va_up, p.neg_inv_gamma), p.cmax) : p.cmax;
    const double mua_up = p.inc - c_up;
    double c, va, mua, util;
    if (mua_up >= 0.0) {
      va = va_up; c = c_up; mua = mua_up;
      util = mdiv(pow(c, p.one_m_gamma), p.one_m_gamma, p.r_one_m_gamma);
    } else {
      const double c_down = va_down > 0 ? ep_min(pow(va_down, p.neg_inv_gamma), p.cmax) : p.cmax;
      const double mua_down = p.inc - c_down;
      if ((mua_down <= 0.0) && p.above_amin) {
        va = va_down; c = c_down; mua = mua_down;
        util = mdiv(pow(c, p.one_m_gamma), p.one_m_gamma, p.r_one_m_gamma);
      } else {
        c = p.inc;
        va = p.va3;
        mua = 0.0;
        util = p.util3;
      }
    }
    out[0] = -(util + mua * va + p.muy * vy + 0.5 * vyy * p.sig2 - p.rho * v);
    if (SAVE) { sv[0] = mua; sv[1] = c; }
  }
};

// Synthetic 3-state HJB of config C5; definition shared verbatim with
// oracle/models.py:hjb3d_synth.
struct Hjb3dSynth : ModelDefaults {
  static constexpr int N = 3, F = 1, NPARAMS = 9, NSAVED = 0;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "hjb3d_synth"; }
  // the body reads the axial neighbours and, through vxy_up only, the (+1,+1,0) / (-1,-1,0) corners
  EP_HD static constexpr bool slot_used(int k) {
    const int o0 = k % 3 - 1, o1 = (k / 3) % 3 - 1, o2 = k / 9 - 1;
    return (o2 == 0 && (o0 == 0 || o1 == 0 || o0 == o1)) || (o0 == 0 && o1 == 0);
  }
  struct Pre { double xyz, mx, my, mz, hx, hy, hz, cxy, rho, theta; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double kx = prm[0], ky = prm[1], kz = prm[2], sx = prm[3], sy = prm[4], sz = prm[5], corr = prm[6];
    const double x = st[0], y = st[1], z = st[2];
    c.xyz = x + y + z;
    c.mx = -kx * (x - 0.5);
    c.my = -ky * (y - 0.5);
    c.mz = -kz * (z - 0.5);
    c.hx = 0.5 * (sx * sx);
    c.hy = 0.5 * (sy * sy);
    c.hz = 0.5 * (sz * sz);
    c.cxy = corr * sx * sy;
    c.rho = prm[7];
    c.theta = prm[8];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<3>* u, double* out, double*) {
    const double v = u[0].v;
    const double vx = (c.mx >= 0) ? u[0].up[0] : u[0].dn[0];
    const double vy = (c.my >= 0) ? u[0].up[1] : u[0].dn[1];
    const double vz = (c.mz >= 0) ? u[0].up[2] : u[0].dn[2];
    const double drift = c.mx * vx + c.my * vy + c.mz * vz;
    const double diff = c.hx * u[0].dd[0] + c.hy * u[0].dd[1] + c.hz * u[0].dd[2];
    const double cross = c.cxy * u[0].xu[0];
    out[0] = -(c.xyz + drift + diff + cross - c.rho * v - c.theta * (v * v));
  }
};

// Synthetic 4-state HJB: what test/runtests.jl:294-320 ("ND sparse coloring": four state variables, 81-point box,
// 16 colors on a 2^4 grid) asks of the stencils beyond the known answer v = x + z + q + r -- upwinded drifts and
// second differences along all four axes and the three cross forms on three different pairs of axes:
// (x, w) vxw_up, (y, z) vyz_down, (x, y) vxy.
struct Hjb4dSynth : ModelDefaults {
  static constexpr int N = 4, F = 1, NPARAMS = 13, NSAVED = 0;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "hjb4d_synth"; }
  struct Pre { double sum, m[4], h[4], cxw, cyz, cxy, rho, theta; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.sum = st[0] + st[1] + st[2] + st[3];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      c.m[d] = -prm[d] * (st[d] - 0.5);
      c.h[d] = 0.5 * (prm[4 + d] * prm[4 + d]);
    }
    c.cxw = prm[8] * prm[4] * prm[7];
    c.cyz = prm[9] * prm[5] * prm[6];
    c.cxy = prm[10] * prm[4] * prm[5];
    c.rho = prm[11];
    c.theta = prm[12];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<4>* u, double* out, double*) {
    const double v = u[0].v;
    double drift = 0.0, diff = 0.0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const double vd = (c.m[d] >= 0) ? u[0].up[d] : u[0].dn[d];
      drift = d == 0 ? c.m[d] * vd : drift + c.m[d] * vd;
      diff = d == 0 ? c.h[d] * u[0].dd[d] : diff + c.h[d] * u[0].dd[d];
    }
    // pairs in the order of `differentiate`: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    const double cross = c.cxw * u[0].xu[2] + c.cyz * u[0].xd[3] + c.cxy * u[0].x[0];
    out[0] = -(c.sum + drift + diff + cross - c.rho * v - c.theta * (v * v));
  }
};

// examples/neoclassical_growth.jl:61-84   saved = (c, muk)
struct NeoclassicalGrowth : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 5, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "neoclassical_growth"; }
  struct Pre { double cap, net, vk3, util3, neg_inv_gamma, one_m_gamma, rho; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double A = prm[0], alpha = prm[1], delta = prm[2], rho = prm[3], gamma = prm[4];
    const double k = st[0];
    c.cap = 10 * A * pow(k, alpha);
    c.net = A * pow(k, alpha) - delta * k;   // A k^alpha - delta k
    c.neg_inv_gamma = -1 / gamma;
    c.one_m_gamma = 1 - gamma;
    c.vk3 = pow(c.net, -gamma);
    c.util3 = pow(c.net, c.one_m_gamma) / c.one_m_gamma;
    c.rho = rho;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& p, const Deriv<1>* u, double* out, double* sv) {
    const double v = u[0].v, vk_up = u[0].up[0], vk_down = u[0].dn[0];
    const double c_up = vk_up >= 0 ? ep_min(pow(vk_up, p.neg_inv_gamma), p.cap) : p.cap;
    const double muk_up = p.net - c_up;
    double c, vk, muk, util;
    if (muk_up > 0) {
      c = c_up; vk = vk_up; muk = muk_up;
      util = pow(c, p.one_m_gamma) / p.one_m_gamma;
    } else {
      const double c_down = vk_down >= 0 ? ep_min(pow(vk_down, p.neg_inv_gamma), p.cap) : p.cap;
      const double muk_down = p.net - c_down;
      if (muk_down < 0) {
        c = c_down; vk = vk_down; muk = muk_down;
        util = pow(c, p.one_m_gamma) / p.one_m_gamma;
      } else {
        muk = 0.0;
        c = p.net;
        vk = p.vk3;
        util = p.util3;
      }
    }
    out[0] = -(util + muk * vk - p.rho * v);
    if (SAVE) { sv[0] = c; sv[1] = muk; }
  }
};

// test/runtests.jl:28-44 (branch structure differs from the example above)
struct GrowthTest : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 5, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "growth_test"; }
  using Pre = NeoclassicalGrowth::Pre;
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) { NeoclassicalGrowth::prepare(prm, st, c); }
  template <bool SAVE>
  EP_HD static void eval(const Pre& p, const Deriv<1>* u, double* out, double* sv) {
    const double vk_up = u[0].up[0], vk_down = u[0].dn[0];
    const double c_up = vk_up > 0 ? ep_min(pow(vk_up, p.neg_inv_gamma), p.cap) : p.cap;
    const double mu_up = p.net - c_up;
    const double c_down = vk_down > 0 ? ep_min(pow(vk_down, p.neg_inv_gamma), p.cap) : p.cap;
    const double mu_down = p.net - c_down;
    double c, vk, muk, util;
    if (mu_up > 0) {
      c = c_up; vk = vk_up; muk = mu_up;
      util = pow(c, p.one_m_gamma) / p.one_m_gamma;
    } else if (mu_down < 0) {
      c = c_down; vk = vk_down; muk = mu_down;
      util = pow(c, p.one_m_gamma) / p.one_m_gamma;
    } else {
      c = p.net;
      vk = p.vk3;
      muk = 0.0;
      util = p.util3;
    }
    out[0] = -(util + muk * vk - p.rho * u[0].v);
    if (SAVE) { sv[0] = c; sv[1] = muk; }
  }
};

// test/runtests.jl:276-279, 314-317:  vt = -(v - x - z - ...)
template <int NDIM>
struct LinearSum : ModelDefaults {
  static constexpr int N = NDIM, F = 1, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() {
    return NDIM == 1 ? "linear_sum_1d" : (NDIM == 2 ? "linear_sum_2d" : (NDIM == 3 ? "linear_sum_3d" : "linear_sum_4d"));
  }
  struct Pre { double st[NDIM]; };
  EP_HD static void prepare(const double*, const double* st, Pre& c) {
#pragma unroll
    for (int d = 0; d < NDIM; ++d) c.st[d] = st[d];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<NDIM>* u, double* out, double*) {
    double acc = u[0].v;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) acc = acc - c.st[d];
    out[0] = -acc;
  }
};

// test/runtests.jl:379-384: two unknowns (v, p): v = x + p, p = v/2 written forward
struct ForwardPrice : ModelDefaults {
  static constexpr int N = 1, F = 2, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "forward_price"; }
  struct Pre { double x; };
  EP_HD static void prepare(const double*, const double* st, Pre& c) { c.x = st[0]; }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    out[0] = -(c.x + u[1].v - u[0].v);
    out[1] = u[0].v / 2 - u[1].v;
  }
};

// test/runtests.jl:227: vt = v - 2 (used with lower_bound for the MCP known answer)
struct VMinusTwo : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 0, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "v_minus_two"; }
  struct Pre { int unused; };
  EP_HD static void prepare(const double*, const double*, Pre& c) { c.unused = 0; }
  template <bool SAVE>
  EP_HD static void eval(const Pre&, const Deriv<1>* u, double* out, double*) {
    out[0] = u[0].v - 2.0;
  }
};

// Two coupled HJB equations on a 2-D grid (regime switching between two drift/payoff regimes with
// intensities lam12, lam21, correlated diffusions, quadratic damping): the synthetic stand-in for
// the reference's multi-unknown 2-D examples (di_tella.jl: 3 unknowns on (x, nu)); exercises F x F
// blocks on the full +-1 box.  Definition shared verbatim with oracle/models.py:hjb2d_regimes.
struct Hjb2dRegimes : ModelDefaults {
  static constexpr int N = 2, F = 2, NPARAMS = 10, NSAVED = 0;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "hjb2d_regimes"; }
  struct Pre { double pay1, pay2, mx1, mx2, my, hx, hy, cxy, rho, lam12, lam21, theta; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double kx1 = prm[0], kx2 = prm[1], ky = prm[2], sx = prm[3], sy = prm[4], corr = prm[5];
    const double x = st[0], y = st[1];
    c.pay1 = x + y;
    c.pay2 = 2 * x - y + 1;
    c.mx1 = -kx1 * (x - 0.5);
    c.mx2 = -kx2 * (x - 0.3);
    c.my = -ky * (y - 0.5);
    c.hx = 0.5 * (sx * sx);
    c.hy = 0.5 * (sy * sy);
    c.cxy = corr * sx * sy;
    c.rho = prm[6];
    c.lam12 = prm[7];
    c.lam21 = prm[8];
    c.theta = prm[9];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<2>* u, double* out, double*) {
    const double v = u[0].v, w = u[1].v;
    const double vx = (c.mx1 >= 0) ? u[0].up[0] : u[0].dn[0];
    const double vy = (c.my >= 0) ? u[0].up[1] : u[0].dn[1];
    const double wx = (c.mx2 >= 0) ? u[1].up[0] : u[1].dn[0];
    const double wy = (c.my >= 0) ? u[1].up[1] : u[1].dn[1];
    out[0] = -(c.pay1 + c.mx1 * vx + c.my * vy + c.hx * u[0].dd[0] + c.hy * u[0].dd[1] + c.cxy * u[0].x[0] -
               c.rho * v + c.lam12 * (w - v) - c.theta * (v * v));
    out[1] = -(c.pay2 + c.mx2 * wx + c.my * wy + c.hx * u[1].dd[0] + c.hy * u[1].dd[1] + c.cxy * u[1].x[0] -
               c.rho * w + c.lam21 * (v - w) - c.theta * (w * w));
  }
};

// examples/asset_pricing/tuckman_vila.jl:67-86  state z, time tau (three-argument form)
struct TuckmanVila : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 6, NSAVED = 0;
  static constexpr int TIME_SLOT = NPARAMS;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "tuckman_vila"; }
  struct Pre { double z, muL, muS, muL_s2, muS_s2, inv_a, a, s2, hs2, rho_z; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& p) {
    const double c = prm[0], r = prm[1], rho = prm[2], sigma = prm[3], a = prm[4], T = prm[5], tau = prm[TIME_SLOT];
    const double z = st[0];
    const double q = 1 + z * z;
    const double phi = z * pow(q, -1 / 2.0);
    const double phiz = pow(q, -3 / 2.0);
    const double phizz = -3 * z * pow(q, -5 / 2.0);
    const double stau = c / r * (1 - exp(-r * (T - tau)));
    const double s2 = sigma * sigma;
    p.muL = rho * z - 0.5 * s2 * phizz / phiz + c / stau * (phi - 1) / phiz;
    p.muS = rho * z - 0.5 * s2 * phizz / phiz + c / stau * (phi + 1) / phiz;
    p.muL_s2 = p.muL / s2;
    p.muS_s2 = p.muS / s2;
    p.inv_a = 1 / a;
    p.a = a;
    p.s2 = s2;
    p.hs2 = 0.5 * s2;
    p.rho_z = rho * z;
    p.z = z;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& p, const Deriv<1>* u, double* out, double*) {
    const double Fzz = u[0].dd[0];
    const double Fz = (p.z >= 0) ? u[0].up[0] : u[0].dn[0];
    const double iL = p.inv_a * (p.muL_s2 + Fz);
    const double iS = p.inv_a * (p.muS_s2 + Fz);
    double mu = 0.0, i = 0.0;
    if (iL > 0) { i = iL; mu = p.muL; }
    else if (iS < 0) { i = iS; mu = p.muS; }
    const double ai = p.a * i;
    out[0] = -((mu + p.s2 * Fz) * p.a * i - p.hs2 * (ai * ai) - p.rho_z * Fz + p.hs2 * (Fzz - Fz * Fz));
  }
};

// examples/asset_pricing/di_tella.jl:92-146   states (x, nu), unknowns (pA, pB, p); p is an algebraic
// (constraint) variable: is_algebraic = (pA = false, pB = false, p = true).
// saved = (sigmax, Q, r, kappa, sigmap, growth, investment)
struct DiTella : ModelDefaults {
  static constexpr int N = 2, F = 3, NPARAMS = 12, NSAVED = 7;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "di_tella"; }
  struct Pre {
    double gamma, psi, rho, tau, A, delta, B, sigma, phi;
    double x, nu, munu, signu;
  };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.gamma = prm[0]; c.psi = prm[1]; c.rho = prm[2]; c.tau = prm[3]; c.A = prm[4]; c.delta = prm[5];
    c.B = prm[6]; c.sigma = prm[7]; c.phi = prm[8];
    const double nubar = prm[9], kappanu = prm[10], signubar = prm[11];
    c.x = st[0];
    c.nu = st[1];
    c.munu = kappanu * (nubar - c.nu);
    c.signu = signubar * sqrt(c.nu);
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<2>* u, double* out, double* sv) {
    const double gamma = c.gamma, psi = c.psi, rho = c.rho, tau = c.tau, A = c.A, delta = c.delta, B = c.B,
                 sigma = c.sigma, phi = c.phi, x = c.x, nu = c.nu, munu = c.munu, signu = c.signu;
    const double pA = u[0].v, pB = u[1].v, p = u[2].v;
    const double q = (p - B) / (2 * A);
    const double g = q - delta;
    const double i = A * (q * q) + B * q;
    const double pAnu = (munu >= 0) ? u[0].up[1] : u[0].dn[1];
    const double pBnu = (munu >= 0) ? u[1].up[1] : u[1].dn[1];
    const double pnu = (munu >= 0) ? u[2].up[1] : u[2].dn[1];
    const double w = x * (1 - x) * (1 - gamma) / (gamma * (psi - 1));   // x (1-x) (1-γ) / (γ (ψ-1))
    const double cg = (1 - gamma) / (gamma * (psi - 1));
    double pAx = u[0].up[0], pBx = u[1].up[0], px = u[2].up[0];
    double sigX, sigpA, sigpB, sigp, kappa, kidio, sigA, nuA, sigB, muX;
    // the two-pass upwinding of the reference (`@goto start`): evaluate with the _up derivatives in x,
    // redo once with the _down ones when the drift of x is not positive
    for (int pass = 0;; ++pass) {
      sigX = w * (pAnu / pA - pBnu / pB) * signu / (1 - w * (pAx / pA - pBx / pB));
      sigpA = pAx / pA * sigX + pAnu / pA * signu;
      sigpB = pBx / pB * sigX + pBnu / pB * signu;
      sigp = px / p * sigX + pnu / p * signu;
      kappa = (sigp + sigma - cg * (x * sigpA + (1 - x) * sigpB)) / (1 / gamma);
      kidio = gamma * phi * nu / x;
      sigA = kappa / gamma + cg * sigpA;
      nuA = kidio / gamma;
      sigB = kappa / gamma + cg * sigpB;
      muX = x * (1 - x) * ((sigA * kappa + nuA * kidio - 1 / pA - tau) - (sigB * kappa - 1 / pB + tau * x / (1 - x)) -
                           (sigA - sigB) * (sigma + sigp));
      if (pass == 0 && muX <= 0) {
        pAx = u[0].dn[0]; pBx = u[1].dn[0]; px = u[2].dn[0];
        continue;
      }
      break;
    }
    const bool cup = sigX * signu >= 0;
    const double pAxnu = cup ? u[0].xu[0] : u[0].xd[0];
    const double pBxnu = cup ? u[1].xu[0] : u[1].xd[0];
    const double pxnu = cup ? u[2].xu[0] : u[2].xd[0];
    const double sX2 = sigX * sigX, sn2 = signu * signu;
    const double mupA = pAx / pA * muX + pAnu / pA * munu + 0.5 * u[0].dd[0] / pA * sX2 + 0.5 * u[0].dd[1] / pA * sn2 +
                        pAxnu / pA * sigX * signu;
    const double mupB = pBx / pB * muX + pBnu / pB * munu + 0.5 * u[1].dd[0] / pB * sX2 + 0.5 * u[1].dd[1] / pB * sn2 +
                        pBxnu / pB * sigX * signu;
    const double mup = px / p * muX + pnu / p * munu + 0.5 * u[2].dd[0] / p * sX2 + 0.5 * u[2].dd[1] / p * sn2 +
                       pxnu / p * sigX * signu;
    const double pn = phi * nu;
    const double r = (1 - i) / p + g + mup + sigma * sigp - kappa * (sigma + sigp) - gamma / x * (pn * pn);
    out[0] = -pA * (1 / pA + (psi - 1) * tau / (1 - gamma) * (pow(pA / pB, (1 - gamma) / (1 - psi)) - 1) - psi * rho +
                    (psi - 1) * (r + kappa * sigA + kidio * nuA) + mupA - (psi - 1) * gamma / 2 * (sigA * sigA + nuA * nuA) +
                    (2 - psi - gamma) / (2 * (psi - 1)) * (sigpA * sigpA) + (1 - gamma) * sigpA * sigA);
    out[1] = -pB * (1 / pB - psi * rho + (psi - 1) * (r + kappa * sigB) + mupB - (psi - 1) * gamma / 2 * (sigB * sigB) +
                    (2 - psi - gamma) / (2 * (psi - 1)) * (sigpB * sigpB) + (1 - gamma) * sigpB * sigB);
    out[2] = -p * ((1 - i) / p - x / pA - (1 - x) / pB);
    if (SAVE) {
      sv[0] = sigX; sv[1] = pow(pA / pB, 1 / (psi - 1)); sv[2] = r; sv[3] = kappa; sv[4] = sigp; sv[5] = g; sv[6] = i;
    }
  }
};

// examples/consumption_saving/consumption_saving_two_income.jl:68-112   state a, unknowns (vl, vh)
// saved = (cl, ch, mula, muha)
struct AhlmTwoStates : ModelDefaults {
  static constexpr int N = 1, F = 2, NPARAMS = 9, NSAVED = 4;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "ahlm_two_states"; }
  struct Pre { double yl, yh, llh, lhl, r, rho, gamma, amin, a, clmax, chmax, incl, inch, nig, omg; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.nig = -1 / prm[6];
    c.omg = 1 - prm[6];
    c.yl = prm[0]; c.yh = prm[1]; c.llh = prm[2]; c.lhl = prm[3]; c.r = prm[4]; c.rho = prm[5]; c.gamma = prm[6];
    c.amin = prm[7];   // prm[8] = amax is only used to build the grid
    c.a = st[0];
    c.clmax = 100.0 * (c.yl + c.r * ep_max(c.a, 0.0));
    c.chmax = 100.0 * (c.yh + c.r * ep_max(c.a, 0.0));
    c.incl = c.yl + c.r * c.a;
    c.inch = c.yh + c.r * c.a;
  }
  // one income state: upwind the asset drift, cap consumption, pin the drift at the borrowing limit
  EP_HD static void branch(const Pre& p, double va_up, double va_down, double cmax, double inc, double& va, double& c,
                           double& mua) {
    const double c_up = va_up > 0 ? ep_min(pow(va_up, p.nig), cmax) : cmax;
    const double mua_up = inc - c_up;
    if (mua_up >= 0.0) {
      va = va_up; c = c_up; mua = mua_up;
    } else {
      const double c_down = va_down > 0 ? ep_min(pow(va_down, p.nig), cmax) : cmax;
      const double mua_down = inc - c_down;
      if (mua_down <= 0.0 && p.a > p.amin) {
        va = va_down; c = c_down; mua = mua_down;
      } else {
        c = inc;
        mua = 0.0;
        va = pow(c, -p.gamma);
      }
    }
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& p, const Deriv<1>* u, double* out, double* sv) {
    const double vl = u[0].v, vh = u[1].v;
    double vla, cl, mula, vha, ch, muha;
    branch(p, u[0].up[0], u[0].dn[0], p.clmax, p.incl, vla, cl, mula);
    out[0] = -(pow(cl, p.omg) / p.omg + mula * vla + p.llh * (vh - vl) - p.rho * vl);
    branch(p, u[1].up[0], u[1].dn[0], p.chmax, p.inch, vha, ch, muha);
    out[1] = -(pow(ch, p.omg) / p.omg + muha * vha + p.lhl * (vl - vh) - p.rho * vh);
    if (SAVE) { sv[0] = cl; sv[1] = ch; sv[2] = mula; sv[3] = muha; }
  }
};

// examples/asset_pricing/garleanu_panageas.jl:74-117   state x, unknowns (pA, pB, phi1, phi2)
// saved = (r, kappa).  B1, B2 arrive normalised as the struct's constructor stores them (:35-39).
struct GarleanuPanageas : ModelDefaults {
  static constexpr int N = 1, F = 4, NPARAMS = 14, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "garleanu_panageas"; }
  struct Pre { double gA, psA, gB, psB, rho, delta, nuA, mu, sigma, B1, d1, B2, d2, omega, x; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.gA = prm[0]; c.psA = prm[1]; c.gB = prm[2]; c.psB = prm[3]; c.rho = prm[4]; c.delta = prm[5]; c.nuA = prm[6];
    c.mu = prm[7]; c.sigma = prm[8]; c.B1 = prm[9]; c.d1 = prm[10]; c.B2 = prm[11]; c.d2 = prm[12]; c.omega = prm[13];
    c.x = st[0];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double* sv) {
    const double gA = c.gA, psA = c.psA, gB = c.gB, psB = c.psB, rho = c.rho, delta = c.delta, nuA = c.nuA, mu = c.mu,
                 sigma = c.sigma, B1 = c.B1, d1 = c.d1, B2 = c.B2, d2 = c.d2, omega = c.omega, x = c.x;
    const double pA = u[0].v, pB = u[1].v, f1 = u[2].v, f2 = u[3].v;
    double pAx = u[0].up[0], pBx = u[1].up[0], f1x = u[2].up[0], f2x = u[3].up[0];
    double G, sigx, spA, spB, sf1, sf2, kappa, sCA, sCB, mcA, mcB, r, muCA, muCB, mux;
    for (int pass = 0;; ++pass) {   // `@goto start`: redo once with the _down derivatives when the drift is not positive
      G = 1 / (x / gA + (1 - x) / gB);
      sigx = sigma * x * (G / gA - 1) /
             (1 + G * x * (1 - x) / (gA * gB) * ((1 - gB * psB) / (psB - 1) * (pBx / pB) - (1 - gA * psA) / (psA - 1) * (pAx / pA)));
      spA = pAx / pA * sigx;
      spB = pBx / pB * sigx;
      sf1 = f1x / f1 * sigx;
      sf2 = f2x / f2 * sigx;
      kappa = G * (sigma - x * (1 - gA * psA) / (gA * (psA - 1)) * spA - (1 - x) * (1 - gB * psB) / (gB * (psB - 1)) * spB);
      sCA = kappa / gA + (1 - gA * psA) / (gA * (psA - 1)) * spA;
      sCB = kappa / gB + (1 - gB * psB) / (gB * (psB - 1)) * spB;
      mcA = kappa * kappa * (1 + psA) / (2 * gA) + (1 - psA * gA) / (gA * (psA - 1)) * kappa * spA -
            (1 - gA * psA) / (2 * gA * (psA - 1)) * (spA * spA);
      mcB = kappa * kappa * (1 + psB) / (2 * gB) + (1 - psB * gB) / (gB * (psB - 1)) * kappa * spB -
            (1 - gB * psB) / (2 * gB * (psB - 1)) * (spB * spB);
      r = rho + 1 / (psA * x + psB * (1 - x)) *
                    (mu - x * mcA - (1 - x) * mcB - delta * ((nuA / pA + (1 - nuA) / pB) * omega * (B1 * f1 + B2 * f2) - 1));
      muCA = psA * (r - rho) + mcA;
      muCB = psB * (r - rho) + mcB;
      mux = x * (muCA - mu) + delta * (nuA / pA * omega * (B1 * f1 + B2 * f2) - x) - sigma * sigx;
      if (pass == 0 && mux <= 0) {
        pAx = u[0].dn[0]; pBx = u[1].dn[0]; f1x = u[2].dn[0]; f2x = u[3].dn[0];
        continue;
      }
      break;
    }
    const double sx2 = sigx * sigx;
    const double mupA = pAx / pA * mux + 0.5 * u[0].dd[0] / pA * sx2;
    const double mupB = pBx / pB * mux + 0.5 * u[1].dd[0] / pB * sx2;
    const double muf1 = f1x / f1 * mux + 0.5 * u[2].dd[0] / f1 * sx2;
    const double muf2 = f2x / f2 * mux + 0.5 * u[3].dd[0] / f2 * sx2;
    out[0] = -pA * (1 / pA + muCA + mupA + sCA * spA - r - delta - kappa * (spA + sCA));
    out[1] = -pB * (1 / pB + muCB + mupB + sCB * spB - r - delta - kappa * (spB + sCB));
    out[2] = -f1 * (1 / f1 + mu - delta - d1 + muf1 + sigma * sf1 - r - kappa * (sf1 + sigma));
    out[3] = -f2 * (1 / f2 + mu - delta - d2 + muf2 + sigma * sf2 - r - kappa * (sf2 + sigma));
    if (SAVE) { sv[0] = r; sv[1] = kappa; }
  }
};

// examples/corporate_finance/leland.jl:44-53   state delta (cash flow), unknown E (equity value);
// solved with lower_bound = 0 (the default option) and bc = (Eδ = (0, 1/(r - μ)))
struct Leland : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 5, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "leland"; }
  struct Pre { double mud, hs2, f, r; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double r = prm[0], mu = prm[1], sigma = prm[2], C = prm[3], tau = prm[4];
    const double d = st[0];
    c.mud = mu * d;
    const double sd = sigma * d;
    c.hs2 = sd * sd;
    c.f = d - (1 - tau) * C;
    c.r = r;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double Ed = (c.mud >= 0) ? u[0].up[0] : u[0].dn[0];
    out[0] = -(c.f + Ed * c.mud + 0.5 * u[0].dd[0] * c.hs2 - c.r * u[0].v);
  }
};

// examples/consumption_saving/wang_wang_yang.jl:50-95   state w (wealth-income ratio), unknown p
// saved = (c, muw); solved with bc = (pw = (1, 1))
struct WangWangYang : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 8, NSAVED = 2;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "wang_wang_yang"; }
  struct Pre { double mu, sigma, r, rho, gamma, psi, wmin, w, mstar, cmax, drift0, s2; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.mu = prm[0]; c.sigma = prm[1]; c.r = prm[2]; c.rho = prm[3]; c.gamma = prm[4]; c.psi = prm[5]; c.wmin = prm[6];
    c.w = st[0];   // prm[7] = wmax only builds the grid
    c.mstar = c.r + c.psi * (c.rho - c.r);
    c.s2 = c.sigma * c.sigma;
    c.drift0 = (c.r - c.mu + c.s2) * c.w;
    c.cmax = 100.0 * (1 + ep_max(c.drift0, 0.0));
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& m, const Deriv<1>* u, double* out, double* sv) {
    const double p = u[0].v, pw_up = u[0].up[0], pw_down = u[0].dn[0], pww = u[0].dd[0];
    const double psi = m.psi, mstar = m.mstar, w = m.w;
    const double p_for_trial = ep_max(p, 2.220446049250313e-16);
    const double c_up = pw_up > 0 ? ep_min(mstar * p * pow(pw_up, -psi), m.cmax) : m.cmax;
    const double muw_up = m.drift0 + 1 - c_up;
    double pw, c, muw;
    if (muw_up >= 0) {
      pw = pw_up > 0 ? pw_up : pow(c_up / (mstar * p_for_trial), -1 / psi);
      c = c_up;
      muw = muw_up;
    } else {
      const double c_down = pw_down > 0 ? ep_min(mstar * p * pow(pw_down, -psi), m.cmax) : m.cmax;
      const double muw_down = m.drift0 + 1 - c_down;
      if (muw_down <= 0 && w > m.wmin) {
        pw = pw_down > 0 ? pw_down : pow(c_down / (mstar * p_for_trial), -1 / psi);
        c = c_down;
        muw = muw_down;
      } else {
        muw = 0.0;
        c = 1 + m.drift0;
        pw = pow(c / (mstar * p_for_trial), -1 / psi);
      }
    }
    out[0] = -((((m.r + psi * (m.rho - m.r)) * pow(pw, 1 - psi) - psi * m.rho) / (psi - 1) + m.mu - m.gamma * m.s2 / 2) * p +
               ((m.r - m.mu + m.gamma * m.s2) * w + 1) * pw + m.s2 * (w * w) / 2 * (pww - m.gamma * (pw * pw) / p));
    if (SAVE) { sv[0] = c; sv[1] = muw; }
  }
};

// test/runtests.jl:326-372 (monotonicity and sign-convention diagnostics): Vt = s (1 + a V + Vx), with the
// first derivative taken forward (dir >= 0) or backward (dir < 0).  wrong_upwind = (-1, +1, down),
// correct_upwind = (-1, +1, up), flipped_sign = (+1, -1, up), correct_sign = (-1, -1, up).
struct LinearUpwindTest : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 3, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "linear_upwind_test"; }
  struct Pre { double s, a, dir; };
  EP_HD static void prepare(const double* prm, const double*, Pre& c) { c.s = prm[0]; c.a = prm[1]; c.dir = prm[2]; }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double vx = c.dir >= 0 ? u[0].up[0] : u[0].dn[0];
    out[0] = c.s * (1.0 + c.a * u[0].v + vx);
  }
};

// examples/asset_pricing/he_krishnamurthy.jl:45-92   state x (specialists' wealth share), unknown pS
// saved = (pS, p, r, kappa, sigmap, muR, sigmaR, alphaI, mux, sigmax, x, px, pxx, sigmaI)
struct HeKrishnamurthy : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 7, NSAVED = 14;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "he_krishnamurthy"; }
  struct Pre { double m, lam, g, sigma, rho, gamma, l, x; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.m = prm[0]; c.lam = prm[1]; c.g = prm[2]; c.sigma = prm[3]; c.rho = prm[4]; c.gamma = prm[5]; c.l = prm[6];
    c.x = st[0];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double* sv) {
    const double m = c.m, lam = c.lam, g = c.g, sigma = c.sigma, rho = c.rho, gamma = c.gamma, l = c.l, x = c.x;
    const double pS = u[0].v, pSxx = u[0].dd[0];
    double pSx = u[0].up[0];
    double p, px, pxx, aI, sigx, spS, sp, sR, kappa, mux;
    for (int pass = 0;; ++pass) {   // `@goto start`
      const double den = x / pS + (1 - x) * rho;
      const double t = 1 / pS - rho - x / (pS * pS) * pSx;
      p = (1 + l) / den;
      px = -(1 + l) / (den * den) * t;
      pxx = -(1 + l) / (den * den) *
                (-1 / (pS * pS) * pSx - 1 / (pS * pS) * pSx - x / (pS * pS) * pSxx + 2 * x / (pS * pS * pS) * (pSx * pSx)) +
            2 * (1 + l) / (den * den * den) * (t * t);
      aI = 1 / (1 - lam * (1 - x));
      if (x * (1 + m) * aI < 1) aI = 1 / (x * (1 + m));
      sigx = x * (aI - 1) * sigma / (1 - x * (aI - 1) * px / p);
      spS = pSx / pS * sigx;
      sp = px / p * sigx;
      sR = sigma + sp;
      kappa = gamma * (aI * sR - spS);
      mux = x * (1 - x) * ((aI - (1 - aI * x) / (1 - x)) * kappa * sR - 1 / pS + rho - (aI - (1 - aI * x) / (1 - x)) * (sR * sR) -
                           l / ((1 - x) * p));
      if (pass == 0 && mux < 0) {
        pSx = u[0].dn[0];
        continue;
      }
      break;
    }
    const double mupS = pSx / pS * mux + 0.5 * pSxx / pS * (sigx * sigx);
    const double mup = px / p * mux + 0.5 * pxx / p * (sigx * sigx);
    const double r = 1 / p + g + mup + sigma * sp - kappa * sR;
    const double sCS = kappa / gamma;
    const double muCS = (r - rho) / gamma + (1 + 1 / gamma) / (2 * gamma) * (kappa * kappa);
    out[0] = -(1 / pS + muCS + mupS + sCS * spS - r - kappa * (sCS + spS));
    if (SAVE) {
      sv[0] = pS; sv[1] = p; sv[2] = r; sv[3] = kappa; sv[4] = sp; sv[5] = r + kappa * sR; sv[6] = sR; sv[7] = aI;
      sv[8] = mux; sv[9] = sigx; sv[10] = x; sv[11] = px; sv[12] = pxx; sv[13] = aI * sR;
    }
  }
};

// examples/asset_pricing/wachter.jl:66-92   state lambda (disaster intensity), unknown p
// prm[11..13] = mgf(Z, 1), mgf(Z, -gamma), mgf(Z, 1 - gamma) of the jump-size distribution Z ~ Normal(muZ, sigmaZ),
// evaluated once by the host mirror (they depend on the parameters only).
struct Wachter : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 14, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "wachter"; }
  struct Pre { double mu, sigma, rho, gamma, psi, lam, mul, sigl, m1, mg, m1g; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double lambar = prm[2], kappal = prm[3], nul = prm[4];
    c.mu = prm[0]; c.sigma = prm[1]; c.rho = prm[7]; c.gamma = prm[8]; c.psi = prm[9];
    c.m1 = prm[11]; c.mg = prm[12]; c.m1g = prm[13];
    c.lam = st[0];
    c.mul = kappal * (lambar - c.lam);
    c.sigl = nul * sqrt(c.lam);
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double mu = c.mu, sigma = c.sigma, rho = c.rho, gamma = c.gamma, psi = c.psi, lam = c.lam;
    const double p = u[0].v, pll = u[0].dd[0];
    const double pl = (c.mul >= 0) ? u[0].up[0] : u[0].dn[0];
    const double mup = pl / p * c.mul + 0.5 * pll / p * (c.sigl * c.sigl);
    const double sp = pl / p * c.sigl;
    const double kZc = gamma * sigma;
    const double kZl = (gamma * psi - 1) / (psi - 1) * sp;
    const double eta = lam * (c.m1 - 1 + c.mg - 1 - (c.m1g - 1));
    const double r = rho + mu / psi - (1 + 1 / psi) / 2 * gamma * (sigma * sigma) - lam * (c.mg - 1) +
                     (1 / psi - gamma) / (1 - gamma) * lam * (c.m1g - 1) - (gamma * psi - 1) / (2 * (psi - 1)) * (sp * sp);
    out[0] = -p * (1 / p + mu + mup + lam * (c.m1 - 1) - r - kZc * sigma - kZl * sp - eta);
  }
};

// examples/asset_pricing/haddad.jl:63-93   states (mu, v), unknown p; endogenous active-capital share
struct Haddad : ModelDefaults {
  static constexpr int N = 2, F = 1, NPARAMS = 11, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "haddad"; }
  struct Pre { double abar, lam, rho, gamma, psi, muc, sigc, mumu, sigmu, muv, sigv; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double mubar = prm[0], vbar = prm[1], kappa_mu = prm[2], nu_mu = prm[3], kappa_v = prm[4], nu_v = prm[5];
    c.abar = prm[6]; c.lam = prm[7]; c.rho = prm[8]; c.gamma = prm[9]; c.psi = prm[10];
    const double mu = st[0], v = st[1];
    c.muc = mu;
    c.sigc = sqrt(v);
    c.mumu = kappa_mu * (mubar - mu);
    c.sigmu = nu_mu * sqrt(v);
    c.muv = kappa_v * (vbar - v);
    c.sigv = nu_v * sqrt(v);
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<2>* u, double* out, double*) {
    const double gamma = c.gamma, psi = c.psi, sigc = c.sigc;
    const double p = u[0].v;
    const double pmu = (c.mumu >= 0) ? u[0].up[0] : u[0].dn[0];
    const double pv = (c.muv >= 0) ? u[0].up[1] : u[0].dn[1];
    const double sZmu = pmu / p * c.sigmu;
    const double sZv = pv / p * c.sigv;
    const double sp2 = sZmu * sZmu + sZv * sZv;
    const double mup = pmu / p * c.mumu + pv / p * c.muv + 0.5 * u[0].dd[0] / p * (c.sigmu * c.sigmu) +
                       0.5 * u[0].dd[1] / p * (c.sigv * c.sigv);
    double astar = c.abar - sqrt(c.lam * c.abar / (0.5 * gamma * (sigc * sigc + sZmu * sZmu + sZv * sZv)));
    astar = astar > 1.0 ? 1.0 : (astar < 0.0 ? 0.0 : astar);   // clamp(astar, 0.0, 1.0)
    const double a1sq = (astar - 1) * (astar - 1);
    out[0] = -p * (1 / p - c.rho + (1 - 1 / psi) * (c.muc - 0.5 * gamma * (sigc * sigc) * (1 - a1sq)) + mup +
                   (0.5 * (1 / psi - gamma) / (1 - 1 / psi) + 0.5 * gamma * (1 - 1 / psi) * a1sq) * sp2);
  }
};

// examples/corporate_finance/bolton_chen_wang.jl:56-72   state w (cash-capital ratio), unknown v (firm value)
// saved = (v, vw, vww, w); solved with bc = (vw = (vw(0), vw(wmax))), the pair an outer root-finder adjusts
struct BoltonChenWang : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 9, NSAVED = 4;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "bolton_chen_wang"; }
  struct Pre { double r, delta, A, sigma, theta, lam, w; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.r = prm[0]; c.delta = prm[1]; c.A = prm[2]; c.sigma = prm[3]; c.theta = prm[4]; c.lam = prm[5];
    c.w = st[0];   // prm[6..8] = (l, gamma, phi) only enter the outer boundary conditions
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double* sv) {
    const double r = c.r, delta = c.delta, A = c.A, sigma = c.sigma, theta = c.theta, lam = c.lam, w = c.w;
    const double v = u[0].v, vww = u[0].dd[0];
    double vw = u[0].up[0];
    double i = 1 / theta * (v / vw - w - 1);
    double muw = (r - lam) * w + A - i - theta * (i * i) / 2 - (i - delta) * w;
    if (muw <= 0) {   // `@goto start` with the backward difference
      vw = u[0].dn[0];
      i = 1 / theta * (v / vw - w - 1);
    }
    out[0] = -((i - delta) * (v - vw * w) + ((r - lam) * w + A - i - theta * (i * i) / 2) * vw + sigma * sigma / 2 * vww - r * v);
    if (SAVE) { sv[0] = v; sv[1] = vw; sv[2] = vww; sv[3] = w; }
  }
};

// examples/asset_pricing/gomez.jl:58-118   state x (entrepreneurs' consumption share), unknown pH
// saved: the 30 equilibrium quantities the example returns, in its order
struct Gomez : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 14, NSAVED = 30;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "gomez"; }
  struct Pre { double gamma, psi, rho, rhoE, aE, eta, delta, phi, piE, g, sigma, lam, tau, x; };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.gamma = prm[0]; c.psi = prm[1]; c.rho = prm[2]; c.rhoE = prm[3]; c.aE = prm[4];   // prm[5] = nu is unused by the body
    c.eta = prm[6]; c.delta = prm[7]; c.phi = prm[8]; c.piE = prm[9]; c.g = prm[10]; c.sigma = prm[11]; c.lam = prm[12];
    c.tau = prm[13];
    c.x = st[0];
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double* sv) {
    const double gamma = c.gamma, psi = c.psi, rho = c.rho, rhoE = c.rhoE, aE = c.aE, eta = c.eta, delta = c.delta,
                 phi = c.phi, piE = c.piE, g = c.g, sigma = c.sigma, lam = c.lam, tau = c.tau, x = c.x;
    const double pH = u[0].v, pHxx = u[0].dd[0];
    const double p = x / rhoE + (1 - x) * pH;
    const double xW = x / rhoE / p;
    const double aEeff = (xW * aE >= 1) ? 1 / xW : aE;
    double pHx = u[0].up[0];
    double px, sigx, spH, sp, sCE, sCH, sWH, kappa, PhiE, PhiH, wealth, wedge, r, muCE, muCH, mux;
    for (int pass = 0;; ++pass) {   // `@goto start`
      px = 1 / rhoE - pH + (1 - x) * pHx;
      sigx = x * (aEeff - 1) * sigma / (1 - x * aEeff * px / p);
      spH = pHx / pH * sigx;
      sp = px / p * sigx;
      sCE = aEeff * (sigma + sp);
      sCH = x < 1 ? (sigma - x * sCE) / (1 - x) : 0.0;
      sWH = sCH + spH;
      kappa = gamma * sCH + (gamma * psi - 1) / (psi - 1) * spH;
      PhiE = aEeff * kappa * (sigma + sp);
      PhiH = kappa * kappa * (1 + psi) / (2 * gamma) + (1 - psi * gamma) / (gamma * (psi - 1)) * kappa * spH -
             (1 - gamma * psi) / (2 * gamma * (psi - 1)) * (spH * spH);
      wealth = (eta + delta + phi + tau) / (eta + delta) * p;
      wedge = (eta + delta) * ((rhoE * piE + (1 - piE) / pH) * wealth - 1);
      const double Psi = x + (1 - x) * psi;
      const double P = (x * rhoE + (1 - x) * psi * rho) / Psi;
      r = P + 1 / Psi * (g - x * PhiE - (1 - x) * PhiH - wedge) + tau;
      muCE = r - tau - rhoE + PhiE;
      muCH = psi * (r - tau - rho) + PhiH;
      mux = x * (muCE - g) + (eta + delta) * (rhoE * piE * wealth - x) - sigma * sigx;
      if (pass == 0 && mux <= 0) {
        pHx = u[0].dn[0];
        continue;
      }
      break;
    }
    const double mupH = pHx / pH * mux + 0.5 * pHxx / pH * (sigx * sigx);
    const double muWH = muCH + mupH + sCH * spH;
    out[0] = -pH * (1 / pH + muWH - r + tau - kappa * sWH);
    if (SAVE) {
      const double pxx = -2 * pHx + (1 - x) * pHxx;
      const double mup = px / p * mux + 0.5 * pxx / p * (sigx * sigx);
      const double sR = sigma + sp;
      const double muR = r + kappa * sR;
      const double sErel = (aEeff - 1) * (sigma + sp);
      const double muErel = muCE - (g + mup + sigma * sp) - sErel * (sigma + sp);
      const double alpha = x < 1 ? (1 - xW * aEeff) / (1 - xW) : 0.0;
      const double sHrel = (alpha - 1) * (sigma + sp);
      const double muHrel = muWH - (g + mup + sigma * sp) - sHrel * (sigma + sp);
      const double muxW = xW * (muCE - (g + mup + sigma * sp) - sErel * (sigma + sp)) +
                          (eta + delta) * (piE * (eta + delta + phi) / (eta + delta) - xW);
      const double sxW = xW * (aEeff - 1) * (sigma + sp);
      sv[0] = x; sv[1] = mux; sv[2] = sigx; sv[3] = kappa; sv[4] = r; sv[5] = sWH; sv[6] = muWH; sv[7] = sp; sv[8] = p;
      sv[9] = px; sv[10] = sR; sv[11] = muR; sv[12] = r + lam * (muR - r); sv[13] = lam * sR; sv[14] = muErel;
      sv[15] = sErel; sv[16] = muHrel; sv[17] = sHrel; sv[18] = muCE; sv[19] = muxW; sv[20] = sxW; sv[21] = pH;
      sv[22] = muCH; sv[23] = wedge; sv[24] = xW; sv[25] = sCE; sv[26] = sCH; sv[27] = aEeff; sv[28] = 1 / rhoE;
      sv[29] = mup;
    }
  }
};

// examples/consumption_saving/consumption_saving_risky_asset.jl:84-192   states (y, a), unknown v
// saved = (v, c, k, va, vaa, vy, y, a, mua); solved with bc = (va = (lower row, upper row)) -- boundary values set for
// one (unknown, state) pair only, which exercises the one-slot-per-point bc layout of src/pdesolve.jl:312-335
struct AhlmRiskyAsset : ModelDefaults {
  static constexpr int N = 2, F = 1, NPARAMS = 10, NSAVED = 9;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "ahlm_risky_asset"; }
  struct Pre { double sigma_y, sigR, rho, gamma, y, a, muy, cash, prem, kmax, cmax; bool at_amax, above_amin; };
  static constexpr double kCmin = 1.4901161193847656e-08;   // sqrt(eps())
  EP_HD static double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    const double kappa_y = prm[0], ybar = prm[1], r = prm[3], muR = prm[4], amin = prm[8], amax = prm[9];
    c.sigma_y = prm[2]; c.sigR = prm[5]; c.rho = prm[6]; c.gamma = prm[7];
    c.y = st[0]; c.a = st[1];
    c.muy = kappa_y * (ybar - c.y);
    c.cash = c.y + r * c.a;
    c.prem = muR - r;
    c.kmax = ep_max(c.a - amin, 0.0);
    c.cmax = ep_max(100.0 * (c.y + r * ep_max(c.a, 0.0)), kCmin);
    c.at_amax = c.a == amax || fabs(c.a - amax) <= kCmin * ep_max(fabs(c.a), fabs(amax));   // a ≈ amax
    c.above_amin = c.a > amin;
  }
  // policy_candidate(va_candidate) of the example
  EP_HD static void candidate(const Pre& m, double vaa, double va_c, double& c_c, double& k_c, double& mua_c) {
    const double sR = m.sigR, prem = m.prem, kmax = m.kmax;
    c_c = va_c > 0.0 ? clampd(pow(va_c, -1 / m.gamma), kCmin, m.cmax) : m.cmax;
    const double H_kmax = prem * va_c * kmax + 0.5 * vaa * (sR * sR) * (kmax * kmax);
    k_c = H_kmax > 0.0 ? kmax : 0.0;
    if (va_c > 0.0 && vaa < 0.0) {
      const double k_merton = clampd(-prem * va_c / (sR * sR * vaa), 0.0, kmax);
      const double H_merton = prem * va_c * k_merton + 0.5 * vaa * (sR * sR) * (k_merton * k_merton);
      const double H_candidate = prem * va_c * k_c + 0.5 * vaa * (sR * sR) * (k_c * k_c);
      k_c = H_merton >= H_candidate ? k_merton : k_c;
    }
    mua_c = m.cash + prem * k_c - c_c;
  }
  template <bool SAVE>
  EP_HD static void eval(const Pre& m, const Deriv<2>* u, double* out, double* sv) {
    const double gamma = m.gamma, sR = m.sigR, prem = m.prem, cash = m.cash, cmin = kCmin;
    const double v = u[0].v, va_up = u[0].up[1], va_down = u[0].dn[1], vyy = u[0].dd[0];
    double vaa = u[0].dd[1];
    if (m.at_amax) {
      const double va_tail = (isfinite(va_down) && va_down > 0.0) ? va_down : va_up;
      vaa = -gamma * va_tail / m.a;
    }
    const double vy = (m.muy >= 0) ? u[0].up[0] : u[0].dn[0];
    double va, c, k, mua;
    candidate(m, vaa, va_up, c, k, mua);
    va = va_up;
    if (!(mua >= 0.0)) {
      double c_down, k_down, mua_down;
      candidate(m, vaa, va_down, c_down, k_down, mua_down);
      if ((mua_down <= 0.0) && m.above_amin) {
        va = va_down; c = c_down; k = k_down; mua = mua_down;
      } else {
        // zero drift: the portfolio solves the first-order condition of the zero-drift Hamiltonian
        double k_hi = m.kmax;
        if (prem < 0.0) k_hi = ep_min(k_hi, ep_max((cash - cmin) / (-prem), 0.0));
        if (k_hi == 0.0 || !(isfinite(cash) && isfinite(prem) && isfinite(vaa) && isfinite(sR))) {
          k = 0.0;
        } else if (vaa < 0.0 && sR > 0.0) {
          const double dh0 = prem * pow(ep_max(cash, cmin), -gamma);
          const double dh1 = prem * pow(ep_max(cash + prem * k_hi, cmin), -gamma) + vaa * (sR * sR) * k_hi;
          if (isfinite(dh0) && isfinite(dh1)) {
            if (dh0 <= 0.0) {
              k = 0.0;
            } else if (dh1 >= 0.0) {
              k = k_hi;
            } else {
              double lo = 0.0, hi = k_hi;
              for (int it = 0; it < 40; ++it) {
                const double mid = (lo + hi) / 2.0;
                const double dh = prem * pow(ep_max(cash + prem * mid, cmin), -gamma) + vaa * (sR * sR) * mid;
                if (dh > 0.0) lo = mid; else hi = mid;
              }
              k = (lo + hi) / 2.0;
            }
          } else {
            k = 0.0;
          }
        } else {
          const double c0 = ep_max(cash, cmin);
          const double c1 = ep_max(cash + prem * k_hi, cmin);
          const double h0 = pow(c0, 1 - gamma) / (1 - gamma);
          const double h1 = pow(c1, 1 - gamma) / (1 - gamma) + 0.5 * vaa * (sR * sR) * (k_hi * k_hi);
          k = h1 > h0 ? k_hi : 0.0;
        }
        c = ep_max(cash + prem * k, cmin);
        va = pow(c, -gamma);
        mua = 0.0;
      }
    }
    const double siga = k * sR;
    out[0] = -(pow(c, 1 - gamma) / (1 - gamma) + va * mua + 0.5 * vaa * (siga * siga) + vy * m.muy +
               0.5 * vyy * (m.sigma_y * m.sigma_y) - m.rho * v);
    if (SAVE) {
      sv[0] = v; sv[1] = c; sv[2] = k; sv[3] = va; sv[4] = vaa; sv[5] = vy; sv[6] = m.y; sv[7] = m.a; sv[8] = mua;
    }
  }
};

// examples/asset_pricing/garleanu_panageas_long_run_risk.jl:74-232 -- the reference's only 3-D example: states
// (x, mu, v); unknowns pA, pB, phi1, phi2 (valuation ratios, backward equations) and r, kC, kM, kV (market-clearing
// prices, forward equations; solved with the signed per-unknown steps of :337-344).  Two-pass upwinding in x
// (`@goto start`, :117-165): the block is evaluated with the _up differences and, if the drift of x comes out <= 0,
// again with the _down ones.  Cross derivatives in the (x, mu) and (x, v) planes only.
// saved = (kappa, r_implied, kC_implied, kM_implied, kV_implied, r_gap, kC_gap, kM_gap, kV_gap, pA_residual,
//          pB_residual, phi1_residual, phi2_residual, mux, sigmax, sigmaxC, sigmaxM, sigmaxV)
struct GarleanuPanageasLRR : ModelDefaults {
  static constexpr int N = 3, F = 8, NPARAMS = 18, NSAVED = 18;
  static constexpr bool USES_CROSS = true;
  static const char* name() { return "garleanu_panageas_lrr"; }
  // axial neighbours and the corners of the (x, mu) and (x, v) planes
  EP_HD static constexpr bool slot_used(int k) {
    const int o1 = (k / 3) % 3 - 1, o2 = k / 9 - 1;
    return o1 == 0 || o2 == 0;
  }
  struct Pre {
    double gA, psA, gB, psB, rho, delta, nuA, B1, d1, B2, d2, omega;
    double x, mu, sc, mumu, smu, muv, sv, G, aA, aB;
  };
  EP_HD static void prepare(const double* prm, const double* st, Pre& c) {
    c.gA = prm[0]; c.psA = prm[1]; c.gB = prm[2]; c.psB = prm[3]; c.rho = prm[4]; c.delta = prm[5]; c.nuA = prm[6];
    const double mubar = prm[7], vbar = prm[8], kmu = prm[9], numu = prm[10], kv = prm[11], nuv = prm[12];
    c.B1 = prm[13]; c.d1 = prm[14]; c.B2 = prm[15]; c.d2 = prm[16]; c.omega = prm[17];
    const double x = st[0], mu = st[1], v = st[2];
    c.x = x; c.mu = mu;
    c.sc = sqrt(v);
    c.mumu = kmu * (mubar - mu);
    c.smu = numu * sqrt(v);
    c.muv = kv * (vbar - v);
    c.sv = nuv * sqrt(v);
    c.G = 1 / (x / c.gA + (1 - x) / c.gB);
    c.aA = (1 - c.gA * c.psA) / (c.gA * (c.psA - 1));
    c.aB = (1 - c.gB * c.psB) / (c.gB * (c.psB - 1));
  }
  struct Pass {
    double pAx_pA, pBx_pB, sxC, sxM, sxV, sx2, spA_C, spA_M, spA_V, spB_C, spB_M, spB_V, sCA_C, sCA_M, sCA_V, sCB_C, sCB_M,
        sCB_V, mcA, mcB, muCA, muCB, mux;
  };
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<3>* u, double* out, double* sv) {
    const double x = c.x, mu = c.mu, sc = c.sc, smu = c.smu, svv = c.sv, gA = c.gA, gB = c.gB, aA = c.aA, aB = c.aB;
    const double pA = u[0].v, pB = u[1].v, f1 = u[2].v, f2 = u[3].v, r = u[4].v, kC = u[5].v, kM = u[6].v, kV = u[7].v;
    const bool upm = c.mumu >= 0, upv = c.muv >= 0;
    const double pAmu = upm ? u[0].up[1] : u[0].dn[1], pBmu = upm ? u[1].up[1] : u[1].dn[1];
    const double f1mu = upm ? u[2].up[1] : u[2].dn[1], f2mu = upm ? u[3].up[1] : u[3].dn[1];
    const double pAv = upv ? u[0].up[2] : u[0].dn[2], pBv = upv ? u[1].up[2] : u[1].dn[2];
    const double f1v = upv ? u[2].up[2] : u[2].dn[2], f2v = upv ? u[3].up[2] : u[3].dn[2];
    const double pAmu_pA = pAmu / pA, pBmu_pB = pBmu / pB, pAv_pA = pAv / pA, pBv_pB = pBv / pB;
    const double k2 = kC * kC + kM * kM + kV * kV;
    const double human_capital = c.omega * (c.B1 * f1 + c.B2 * f2);
    auto block = [&](double pAx, double pBx, Pass& q) {
      q.pAx_pA = pAx / pA;
      q.pBx_pB = pBx / pB;
      const double denom = 1 - x * aA * q.pAx_pA;
      q.sxC = x * (kC / gA - sc) / denom;
      q.sxM = x * (kM / gA + aA * pAmu_pA * smu) / denom;
      q.sxV = x * (kV / gA + aA * pAv_pA * svv) / denom;
      q.sx2 = q.sxC * q.sxC + q.sxM * q.sxM + q.sxV * q.sxV;
      q.spA_C = q.pAx_pA * q.sxC;
      q.spA_M = q.pAx_pA * q.sxM + pAmu_pA * smu;
      q.spA_V = q.pAx_pA * q.sxV + pAv_pA * svv;
      q.spB_C = q.pBx_pB * q.sxC;
      q.spB_M = q.pBx_pB * q.sxM + pBmu_pB * smu;
      q.spB_V = q.pBx_pB * q.sxV + pBv_pB * svv;
      q.sCA_C = kC / gA + aA * q.spA_C;
      q.sCA_M = kM / gA + aA * q.spA_M;
      q.sCA_V = kV / gA + aA * q.spA_V;
      q.sCB_C = kC / gB + aB * q.spB_C;
      q.sCB_M = kM / gB + aB * q.spB_M;
      q.sCB_V = kV / gB + aB * q.spB_V;
      const double spA2 = q.spA_C * q.spA_C + q.spA_M * q.spA_M + q.spA_V * q.spA_V;
      const double spB2 = q.spB_C * q.spB_C + q.spB_M * q.spB_M + q.spB_V * q.spB_V;
      const double kspA = kC * q.spA_C + kM * q.spA_M + kV * q.spA_V;
      const double kspB = kC * q.spB_C + kM * q.spB_M + kV * q.spB_V;
      q.mcA = (1 + c.psA) / (2 * gA) * k2 + aA * kspA - 0.5 * aA * spA2;
      q.mcB = (1 + c.psB) / (2 * gB) * k2 + aB * kspB - 0.5 * aB * spB2;
      q.muCA = c.psA * (r - c.rho) + q.mcA;
      q.muCB = c.psB * (r - c.rho) + q.mcB;
      q.mux = x * (q.muCA - mu) + c.delta * (c.nuA / pA * human_capital - x) - sc * q.sxC;
    };
    Pass q;
    block(u[0].up[0], u[1].up[0], q);
    double f1x = u[2].up[0], f2x = u[3].up[0];
    if (q.mux <= 0) {
      block(u[0].dn[0], u[1].dn[0], q);
      f1x = u[2].dn[0];
      f2x = u[3].dn[0];
    }
    const bool cm = q.sxM * smu >= 0, cv = q.sxV * svv >= 0;
    // pairs: (x, mu) = 0, (x, v) = 1
    const double pAxmu = cm ? u[0].xu[0] : u[0].xd[0], pBxmu = cm ? u[1].xu[0] : u[1].xd[0];
    const double f1xmu = cm ? u[2].xu[0] : u[2].xd[0], f2xmu = cm ? u[3].xu[0] : u[3].xd[0];
    const double pAxv = cv ? u[0].xu[1] : u[0].xd[1], pBxv = cv ? u[1].xu[1] : u[1].xd[1];
    const double f1xv = cv ? u[2].xu[1] : u[2].xd[1], f2xv = cv ? u[3].xu[1] : u[3].xd[1];
    const double f1x_f1 = f1x / f1, f2x_f2 = f2x / f2, f1mu_f1 = f1mu / f1, f2mu_f2 = f2mu / f2;
    const double f1v_f1 = f1v / f1, f2v_f2 = f2v / f2;
    const double sf1_C = f1x_f1 * q.sxC, sf1_M = f1x_f1 * q.sxM + f1mu_f1 * smu, sf1_V = f1x_f1 * q.sxV + f1v_f1 * svv;
    const double sf2_C = f2x_f2 * q.sxC, sf2_M = f2x_f2 * q.sxM + f2mu_f2 * smu, sf2_V = f2x_f2 * q.sxV + f2v_f2 * svv;
    const double smu2 = smu * smu, sv2 = svv * svv;
    const double mupA = q.pAx_pA * q.mux + pAmu_pA * c.mumu + pAv_pA * c.muv + 0.5 * u[0].dd[0] / pA * q.sx2 +
                        0.5 * u[0].dd[1] / pA * smu2 + 0.5 * u[0].dd[2] / pA * sv2 + pAxmu / pA * q.sxM * smu +
                        pAxv / pA * q.sxV * svv;
    const double mupB = q.pBx_pB * q.mux + pBmu_pB * c.mumu + pBv_pB * c.muv + 0.5 * u[1].dd[0] / pB * q.sx2 +
                        0.5 * u[1].dd[1] / pB * smu2 + 0.5 * u[1].dd[2] / pB * sv2 + pBxmu / pB * q.sxM * smu +
                        pBxv / pB * q.sxV * svv;
    const double muf1 = f1x_f1 * q.mux + f1mu_f1 * c.mumu + f1v_f1 * c.muv + 0.5 * u[2].dd[0] / f1 * q.sx2 +
                        0.5 * u[2].dd[1] / f1 * smu2 + 0.5 * u[2].dd[2] / f1 * sv2 + f1xmu / f1 * q.sxM * smu +
                        f1xv / f1 * q.sxV * svv;
    const double muf2 = f2x_f2 * q.mux + f2mu_f2 * c.mumu + f2v_f2 * c.muv + 0.5 * u[3].dd[0] / f2 * q.sx2 +
                        0.5 * u[3].dd[1] / f2 * smu2 + 0.5 * u[3].dd[2] / f2 * sv2 + f2xmu / f2 * q.sxM * smu +
                        f2xv / f2 * q.sxV * svv;
    const double sCAspA = q.sCA_C * q.spA_C + q.sCA_M * q.spA_M + q.sCA_V * q.spA_V;
    const double sCBspB = q.sCB_C * q.spB_C + q.sCB_M * q.spB_M + q.sCB_V * q.spB_V;
    const double kspA_CA = kC * (q.spA_C + q.sCA_C) + kM * (q.spA_M + q.sCA_M) + kV * (q.spA_V + q.sCA_V);
    const double kspB_CB = kC * (q.spB_C + q.sCB_C) + kM * (q.spB_M + q.sCB_M) + kV * (q.spB_V + q.sCB_V);
    const double pAt = -pA * (1 / pA + q.muCA + mupA + sCAspA - r - c.delta - kspA_CA);
    const double pBt = -pB * (1 / pB + q.muCB + mupB + sCBspB - r - c.delta - kspB_CB);
    const double ksf1_Y = kC * (sf1_C + sc) + kM * sf1_M + kV * sf1_V;
    const double ksf2_Y = kC * (sf2_C + sc) + kM * sf2_M + kV * sf2_V;
    const double f1t = -f1 * (1 / f1 + mu - c.delta - c.d1 + muf1 + sc * sf1_C - r - ksf1_Y);
    const double f2t = -f2 * (1 / f2 + mu - c.delta - c.d2 + muf2 + sc * sf2_C - r - ksf2_Y);
    const double kC_imp = c.G * (sc - x * aA * q.spA_C - (1 - x) * aB * q.spB_C);
    const double kM_imp = c.G * (0.0 - x * aA * q.spA_M - (1 - x) * aB * q.spB_M);
    const double kV_imp = c.G * (0.0 - x * aA * q.spA_V - (1 - x) * aB * q.spB_V);
    const double r_imp = c.rho + (mu - x * q.mcA - (1 - x) * q.mcB -
                                  c.delta * ((c.nuA / pA + (1 - c.nuA) / pB) * human_capital - 1)) /
                                     (c.psA * x + c.psB * (1 - x));
    out[0] = pAt; out[1] = pBt; out[2] = f1t; out[3] = f2t;
    out[4] = r_imp - r; out[5] = kC_imp - kC; out[6] = kM_imp - kM; out[7] = kV_imp - kV;
    if (SAVE) {
      sv[0] = sqrt(k2); sv[1] = r_imp; sv[2] = kC_imp; sv[3] = kM_imp; sv[4] = kV_imp;
      sv[5] = out[4]; sv[6] = out[5]; sv[7] = out[6]; sv[8] = out[7];
      sv[9] = pAt; sv[10] = pBt; sv[11] = f1t; sv[12] = f2t;
      sv[13] = q.mux; sv[14] = sqrt(q.sx2); sv[15] = q.sxC; sv[16] = q.sxM; sv[17] = q.sxV;
    }
  }
};

// test/runtests.jl:422-431 (residual-growth guard): vt = a v - k v^3 at every grid point (the reference's
// scalar hump!(ydot, y) = y - y^3 with a = k = 1); no derivative enters, so the Jacobian is diagonal.
struct PointwiseCubic : ModelDefaults {
  static constexpr int N = 1, F = 1, NPARAMS = 2, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "pointwise_cubic"; }
  struct Pre { double a, k; };
  EP_HD static void prepare(const double* prm, const double*, Pre& c) { c.a = prm[0]; c.k = prm[1]; }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    const double v = u[0].v;
    out[0] = c.a * v - c.k * (v * v * v);
  }
};

// test/runtests.jl:403-411 (mixed_residual!: ydot = (y1 - 2, 3 - y2), Δ = [1, -1]) and :456-489 (residual!:
// ydot = (2 y1 - 4, y2 - 3); one implicit step with an algebraic second row gives (4/3, 3); bounds give (1, 3)):
// two uncoupled affine rows vt = a1 v + b1, pt = a2 p + b2 at every grid point.
struct AffinePair : ModelDefaults {
  static constexpr int N = 1, F = 2, NPARAMS = 4, NSAVED = 0;
  static constexpr bool USES_CROSS = false;
  static const char* name() { return "affine_pair"; }
  struct Pre { double a1, b1, a2, b2; };
  EP_HD static void prepare(const double* prm, const double*, Pre& c) { c.a1 = prm[0]; c.b1 = prm[1]; c.a2 = prm[2]; c.b2 = prm[3]; }
  template <bool SAVE>
  EP_HD static void eval(const Pre& c, const Deriv<1>* u, double* out, double*) {
    out[0] = c.a1 * u[0].v + c.b1;
    out[1] = c.a2 * u[1].v + c.b2;
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
  X(VMinusTwo)                     \
  X(Hjb2dRegimes)                  \
  X(TuckmanVila)                   \
  X(DiTella)                       \
  X(AhlmTwoStates)                 \
  X(GarleanuPanageas)              \
  X(Leland)                        \
  X(WangWangYang)                  \
  X(LinearUpwindTest)              \
  X(HeKrishnamurthy)               \
  X(Wachter)                       \
  X(Haddad)                        \
  X(BoltonChenWang)                \
  X(Gomez)                         \
  X(AhlmRiskyAsset)                \
  X(PointwiseCubic)                \
  X(AffinePair)                    \
  X(GarleanuPanageasLRR)           \
  X(LinearSum<4>)                  \
  X(Hjb4dSynth)

}  // namespace econpdes
