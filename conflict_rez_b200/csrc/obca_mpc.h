// obca_mpc.h -- distributed-MPC NLP of VehicleFollower.setup_controller (confrez/control/vehicle_follower.py:146-368):
// horizon of N nodes, RK4 x 4 dynamics (confrez/control/dynamic_model.py:30-58), obstacle OBCA triples per node,
// one pair block per other vehicle whose predicted pose is a parameter, tracking cost.
//
// The flat layout is the collocation layout `Lay` with V = 1 vehicle, Mv = N nodes, P = number of others; the YCOL field
// holds the multipliers of the dynamics rows  z_{n+1} - F(z_n, u_n) = 0  (n < N-1), YINIT the five initial-state rows.
// Per instance parameters (stored behind the iterate): cur[5], ref[N][3], others[P][N][3].
//
//   [EVAL]     residuals, Lagrangian gradient (dynamics Jacobian by first-order forward mode)
//   [LOCAL]    obstacle / pair blocks: the same eliminations as the collocation problem (obca_kkt.h)
//   [RICCATI]  classical stage-wise recursion, state 5, control 2; the Lagrangian Hessian of the RK4 map comes from
//              second-order forward-mode jets
#pragma once

namespace obca {

// ------------------------------------------------------------------------------------------------
// forward-mode jets in the 7 inputs (z, u): value, gradient, (optionally) packed Hessian
// ------------------------------------------------------------------------------------------------
template <bool HESS>
struct Jet {
  double v, g[7], h[HESS ? 28 : 1];
};

template <bool H>
OBCA_HD Jet<H> jet_const(double c) {
  Jet<H> r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < 7; ++i) r.g[i] = 0;
  if (H)
#pragma unroll
    for (int i = 0; i < 28; ++i) r.h[i] = 0;
  return r;
}

template <bool H>
OBCA_HD Jet<H> jet_var(double c, int k) {
  Jet<H> r = jet_const<H>(c);
  r.g[k] = 1.0;
  return r;
}

// r = a + s * b
template <bool H>
OBCA_HD Jet<H> jet_axpy(const Jet<H>& a, double s, const Jet<H>& b) {
  Jet<H> r;
  r.v = a.v + s * b.v;
#pragma unroll
  for (int i = 0; i < 7; ++i) r.g[i] = a.g[i] + s * b.g[i];
  if (H)
#pragma unroll
    for (int i = 0; i < 28; ++i) r.h[i] = a.h[i] + s * b.h[i];
  return r;
}

template <bool H>
OBCA_HD Jet<H> jet_mul(const Jet<H>& a, const Jet<H>& b) {
  Jet<H> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < 7; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  if (H)
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) r.h[sym(i, j)] = a.v * b.h[sym(i, j)] + b.v * a.h[sym(i, j)] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}

// r = f(a) given f, f', f'' at a.v
template <bool H>
OBCA_HD Jet<H> jet_unary(const Jet<H>& a, double f0, double f1, double f2) {
  Jet<H> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < 7; ++i) r.g[i] = f1 * a.g[i];
  if (H)
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) r.h[sym(i, j)] = f1 * a.h[sym(i, j)] + f2 * a.g[i] * a.g[j];
  return r;
}

// kinematic bicycle f(z, u) on jets (dynamic_model.py:5-27); z = (x, y, psi, v, delta), u = (a, w)
template <bool H>
OBCA_HD void jet_f(const Jet<H>* z, const Jet<H>* u, double wb, Jet<H>* out) {
  const double c = cos(z[2].v), s = sin(z[2].v), t = tan(z[4].v), sec2 = 1.0 + t * t;
  Jet<H> jc = jet_unary<H>(z[2], c, -s, -c), js = jet_unary<H>(z[2], s, c, -s);
  Jet<H> jt = jet_unary<H>(z[4], t / wb, sec2 / wb, 2.0 * t * sec2 / wb);
  out[0] = jet_mul<H>(z[3], jc);
  out[1] = jet_mul<H>(z[3], js);
  out[2] = jet_mul<H>(z[3], jt);
  out[3] = u[0];
  out[4] = u[1];
}

// RK4 with 4 sub-steps of h = dt / 4 (dynamic_model.py:30-58) on jets
template <bool H>
OBCA_HD void jet_rk4(const double* zu, double dt, double wb, Jet<H>* z) {
  Jet<H> u[2] = {jet_var<H>(zu[5], 5), jet_var<H>(zu[6], 6)};
#pragma unroll
  for (int q = 0; q < 5; ++q) z[q] = jet_var<H>(zu[q], q);
  const double h = dt / 4.0;
  for (int sub = 0; sub < 4; ++sub) {
    Jet<H> a1[5], a2[5], a3[5], a4[5], tmp[5];
    jet_f<H>(z, u, wb, a1);
    for (int q = 0; q < 5; ++q) tmp[q] = jet_axpy<H>(z[q], h / 2, a1[q]);
    jet_f<H>(tmp, u, wb, a2);
    for (int q = 0; q < 5; ++q) tmp[q] = jet_axpy<H>(z[q], h / 2, a2[q]);
    jet_f<H>(tmp, u, wb, a3);
    for (int q = 0; q < 5; ++q) tmp[q] = jet_axpy<H>(z[q], h, a3[q]);
    jet_f<H>(tmp, u, wb, a4);
    for (int q = 0; q < 5; ++q) {
      Jet<H> acc = jet_axpy<H>(z[q], h / 6, a1[q]);
      acc = jet_axpy<H>(acc, h / 3, a2[q]);
      acc = jet_axpy<H>(acc, h / 3, a3[q]);
      z[q] = jet_axpy<H>(acc, h / 6, a4[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// univariate Taylor jets v + a t + b t^2 along one direction d of the 7 inputs: the work of the dynamics derivatives is
// split into (node, direction) tasks that run on different threads.  d'(grad^2 F_r) d = 2 b_r, (dF_r/d(z,u)) d = a_r.
// ------------------------------------------------------------------------------------------------
template <bool SECOND>
struct Tay {
  double v, a, b;
};
template <bool S2>
OBCA_HD Tay<S2> tay_axpy(const Tay<S2>& x, double s, const Tay<S2>& y) {
  Tay<S2> r;
  r.v = x.v + s * y.v, r.a = x.a + s * y.a, r.b = S2 ? x.b + s * y.b : 0.0;
  return r;
}
template <bool S2>
OBCA_HD Tay<S2> tay_mul(const Tay<S2>& x, const Tay<S2>& y) {
  Tay<S2> r;
  r.v = x.v * y.v, r.a = x.v * y.a + x.a * y.v, r.b = S2 ? x.v * y.b + x.a * y.a + x.b * y.v : 0.0;
  return r;
}
template <bool S2>
OBCA_HD Tay<S2> tay_unary(const Tay<S2>& x, double f0, double f1, double f2) {
  Tay<S2> r;
  r.v = f0, r.a = f1 * x.a, r.b = S2 ? f1 * x.b + 0.5 * f2 * x.a * x.a : 0.0;
  return r;
}
template <bool S2>
OBCA_HD void tay_f(const Tay<S2>* z, const Tay<S2>* u, double wb, Tay<S2>* out) {
  const double c = cos(z[2].v), s = sin(z[2].v), t = tan(z[4].v), sec2 = 1.0 + t * t;
  Tay<S2> jc = tay_unary<S2>(z[2], c, -s, -c), js = tay_unary<S2>(z[2], s, c, -s);
  Tay<S2> jt = tay_unary<S2>(z[4], t / wb, sec2 / wb, 2.0 * t * sec2 / wb);
  out[0] = tay_mul<S2>(z[3], jc);
  out[1] = tay_mul<S2>(z[3], js);
  out[2] = tay_mul<S2>(z[3], jt);
  out[3] = u[0];
  out[4] = u[1];
}
// RK4 x 4 (dynamic_model.py:30-58) along z(t) = zu + t d
template <bool S2>
OBCA_HD void tay_rk4(const double* zu, const double* d, double dt, double wb, Tay<S2>* z) {
  Tay<S2> u[2] = {{zu[5], d[5], 0.0}, {zu[6], d[6], 0.0}};
  for (int q = 0; q < 5; ++q) z[q].v = zu[q], z[q].a = d[q], z[q].b = 0.0;
  const double h = dt / 4.0;
  for (int sub = 0; sub < 4; ++sub) {
    Tay<S2> a1[5], a2[5], a3[5], a4[5], tmp[5];
    tay_f<S2>(z, u, wb, a1);
    for (int q = 0; q < 5; ++q) tmp[q] = tay_axpy<S2>(z[q], h / 2, a1[q]);
    tay_f<S2>(tmp, u, wb, a2);
    for (int q = 0; q < 5; ++q) tmp[q] = tay_axpy<S2>(z[q], h / 2, a2[q]);
    tay_f<S2>(tmp, u, wb, a3);
    for (int q = 0; q < 5; ++q) tmp[q] = tay_axpy<S2>(z[q], h, a3[q]);
    tay_f<S2>(tmp, u, wb, a4);
    for (int q = 0; q < 5; ++q) {
      Tay<S2> acc = tay_axpy<S2>(z[q], h / 6, a1[q]);
      acc = tay_axpy<S2>(acc, h / 3, a2[q]);
      acc = tay_axpy<S2>(acc, h / 3, a3[q]);
      z[q] = tay_axpy<S2>(acc, h / 6, a4[q]);
    }
  }
}

// plain RK4 (values only) for the line-search trial points
OBCA_HD void rk4_value(const double* zu, double dt, double wb, double* out) {
  double z[5] = {zu[0], zu[1], zu[2], zu[3], zu[4]};
  const double ua = zu[5], uw = zu[6], h = dt / 4.0;
  for (int sub = 0; sub < 4; ++sub) {
    double k[4][5], zz[5];
    for (int st = 0; st < 4; ++st) {
      const double f = st == 0 ? 0.0 : (st == 3 ? h : h / 2);
      for (int q = 0; q < 5; ++q) zz[q] = st == 0 ? z[q] : z[q] + f * k[st - 1][q];
      k[st][0] = zz[3] * cos(zz[2]);
      k[st][1] = zz[3] * sin(zz[2]);
      k[st][2] = zz[3] * tan(zz[4]) / wb;
      k[st][3] = ua;
      k[st][4] = uw;
    }
    for (int q = 0; q < 5; ++q) z[q] += h / 6 * (k[0][q] + 2 * k[1][q] + 2 * k[2][q] + k[3][q]);
  }
  for (int q = 0; q < 5; ++q) out[q] = z[q];
}

// forward Euler z + dt f(z, u) (Vehicle.state_ws, vehicle.py:169-173) along z(t) = zu + t d, and its plain value
template <bool S2>
OBCA_HD void tay_euler(const double* zu, const double* d, double dt, double wb, Tay<S2>* z) {
  Tay<S2> u[2] = {{zu[5], d[5], 0.0}, {zu[6], d[6], 0.0}}, z0[5], f[5];
  for (int q = 0; q < 5; ++q) z0[q].v = zu[q], z0[q].a = d[q], z0[q].b = 0.0;
  tay_f<S2>(z0, u, wb, f);
  for (int q = 0; q < 5; ++q) z[q] = tay_axpy<S2>(z0[q], dt, f[q]);
}
OBCA_HD void euler_value(const double* zu, double dt, double wb, double* out) {
  out[0] = zu[0] + dt * zu[3] * cos(zu[2]);
  out[1] = zu[1] + dt * zu[3] * sin(zu[2]);
  out[2] = zu[2] + dt * zu[3] * tan(zu[4]) / wb;
  out[3] = zu[3] + dt * zu[5];
  out[4] = zu[4] + dt * zu[6];
}
// per-instance MPC parameters behind the iterate: cur[5], ref[N][3], others[P][N][3]
struct MpcPar {
  const double *cur, *ref, *others;
};
OBCA_HD MpcPar mpc_par(const Lay& L, const Scratch& W) {
  MpcPar p;
  p.cur = W.init_pose;
  p.ref = p.cur + 5;
  p.others = p.ref + 3 * L.Mv;
  return p;
}
inline size_t mpc_param_doubles(const Lay& L) { return 5 + (size_t)3 * L.Mv * (1 + L.P); }

OBCA_HD void load_other_pose(const Lay& L, const MpcPar& par, int o, int n, Pose& p) {
  const double* q = par.others + ((size_t)o * L.Mv + n) * 3;
  p.x = q[0], p.y = q[1], p.psi = q[2];
  p.c = cos(p.psi), p.s = sin(p.psi);
}

// ------------------------------------------------------------------------------------------------
// [EVAL]
// ------------------------------------------------------------------------------------------------
OBCA_HDN void mpc_eval_all(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, const double* x, const double* y, double* c,
                           double* gl, double* f_out, double* gdt_out) {
  OBCA_ASSUME_STATIC(L, S);
  assume_scratch(W);
  const MpcPar par = mpc_par(L, W);
  const int N = L.Mv;
  const bool EU = L.euler != 0;  // state warm start: Euler dynamics, cost a^2 + w^2, tube sets, a_0 = w_0 = 0, final heading
  double* PG = W.PG;
  prof_mark(ctx, 11);
  // task families in units of 32 tasks of one family, longest first, from a shared counter (see mpc_kkt_solve):
  //   pair blocks (other vehicle's pose is a parameter): residuals, gl of the pair variables, own-pose gradient -> PG
  //   (node, obstacle): residual rows, gradient of the block variables, pose gradient -> OG
  //   (node, nonlinear input i): column i of the dynamics Jacobian contracted with y -> JG, residual rows; values only without y
  double f_part = 0;
  double* JG = W.HN;  // scratch (the collocation-mode node buffers are free in MPC mode): [N][5], then [N][O][3]
  double* OG = W.HN + (size_t)N * 5;
  const int nPair = L.P * N, nObs = N * L.O, nDyn = y ? (N - 1) * 5 : N - 1;
  const int uPair = (nPair + 31) / 32, uObs = (nObs + 31) / 32, uDyn = (nDyn + 31) / 32;
  int* next_unit = (int*)(ctx.red + 36);  // shared-memory slot of the reduction scratch that the reductions never touch
  if (ctx.tid == 0) *next_unit = 0;
  cta_sync(ctx);
  for (;;) {
    int unit;
#if defined(__CUDA_ARCH__)
    {
      int u0 = 0;
      if ((ctx.tid & 31) == 0) u0 = atomicAdd(next_unit, 1);
      unit = __shfl_sync(0xffffffffu, u0, 0);
    }
    const int lane0 = ctx.tid & 31, lane1 = lane0 + 1;
#else
    unit = (*next_unit)++;
    const int lane0 = 0, lane1 = 32;
#endif
    if (unit >= uPair + uObs + uDyn) break;
    for (int ln = lane0; ln < lane1; ++ln) {
      if (unit < uPair) {
        const int it = unit * 32 + ln;
        if (it >= nPair) continue;
        int o = it / N, n = it % N;
        Pose a, b;
        load_pose(L, x, 0, n, a);
        load_other_pose(L, par, o, n, b);
        PairBlk B;
        load_pair(L, x, o, n, B);
        pair_residual(S, a, b, B);
        for (int r = 0; r < 6; ++r) c[L.YPAIR(o, r, n)] = B.c[r];
        if (!y) continue;
        double yd = y[L.YPAIR(o, 0, n)], ye1[2] = {y[L.YPAIR(o, 1, n)], y[L.YPAIR(o, 2, n)]};
        double ye2[2] = {y[L.YPAIR(o, 3, n)], y[L.YPAIR(o, 4, n)]}, yn = y[L.YPAIR(o, 5, n)];
        double Rtea[2] = {a.c * ye1[0] + a.s * ye1[1], -a.s * ye1[0] + a.c * ye1[1]};
        double Rteb[2] = {b.c * ye2[0] + b.s * ye2[1], -b.s * ye2[0] + b.c * ye2[1]};
        for (int r = 0; r < 4; ++r) {
          gl[L.PL(o, r, n)] = -yd * B.ba[r] + S.G[r][0] * Rtea[0] + S.G[r][1] * Rtea[1];
          gl[L.PM(o, r, n)] = -yd * B.bb[r] + S.G[r][0] * Rteb[0] + S.G[r][1] * Rteb[1];
        }
        gl[L.PS(o, 0, n)] = ye1[0] - ye2[0] - 2.0 * yn * B.s[0];
        gl[L.PS(o, 1, n)] = ye1[1] - ye2[1] - 2.0 * yn * B.s[1];
        gl[L.PSD(o, n)] = -yd;
        gl[L.PSN(o, n)] = -yn;
        gl[L.PEL(o, n)] = S.rho + yd;
        double dRua[2] = {-a.s * B.ua[0] - a.c * B.ua[1], a.c * B.ua[0] - a.s * B.ua[1]};
        double* g = PG + (size_t)o * 6 * L.Mv + n;
        g[0] = -yd * B.Rua[0];
        g[L.Mv] = -yd * B.Rua[1];
        g[2 * L.Mv] = -yd * (a.x * dRua[0] + a.y * dRua[1]) + ye1[0] * dRua[0] + ye1[1] * dRua[1];
      } else if (unit < uPair + uObs) {
        const int it = (unit - uPair) * 32 + ln;
        if (it >= nObs) continue;
        const int n = it / L.O, j = it % L.O;
        Pose p;
        load_pose(L, x, 0, n, p);
        ObsBlk B;
        for (int r = 0; r < 4; ++r) B.lam[r] = x[L.LAM(0, j, r, n)], B.mu[r] = x[L.MU(0, j, r, n)];
        B.sd = x[L.SD(0, j, n)];
        B.el = x[L.EL(0, j, n)];
        f_part += S.rho * B.el;
        obs_residual(S, j, p, B);
        for (int r = 0; r < 4; ++r) c[L.YOBS(0, j, r, n)] = B.c[r];
        if (!y) continue;
        double y1 = y[L.YOBS(0, j, 0, n)], y2[2] = {y[L.YOBS(0, j, 1, n)], y[L.YOBS(0, j, 2, n)]}, y3 = y[L.YOBS(0, j, 3, n)];
        double Ry[2] = {p.c * y2[0] - p.s * y2[1], p.s * y2[0] + p.c * y2[1]};
        for (int r = 0; r < 4; ++r) {
          const double* A = S.obsA[j][r];
          gl[L.LAM(0, j, r, n)] = y1 * B.Atb[r] + A[0] * Ry[0] + A[1] * Ry[1] + 2.0 * y3 * (A[0] * B.u[0] + A[1] * B.u[1]);
          gl[L.MU(0, j, r, n)] = -y1 * S.g[r] + S.G[r][0] * y2[0] + S.G[r][1] * y2[1];
        }
        gl[L.SD(0, j, n)] = -y1;
        gl[L.EL(0, j, n)] = S.rho + y1;
        double* og = OG + (size_t)it * 3;
        og[0] = y1 * B.u[0];
        og[1] = y1 * B.u[1];
        og[2] = y2[0] * (-p.s * B.u[0] + p.c * B.u[1]) + y2[1] * (-p.c * B.u[0] - p.s * B.u[1]);
      } else if (y) {
        const int it = (unit - uPair - uObs) * 32 + ln;
        if (it >= nDyn) continue;
        const int n = it / 5, i = it % 5;
        double z[NZ], d[NZ] = {0, 0, 0, 0, 0, 0, 0};
        for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(0, q, n)];
        d[2 + i] = 1.0;
        Tay<false> F[5];
        if (EU) tay_euler<false>(z, d, S.dt_mpc, S.wb, F);
        else tay_rk4<false>(z, d, S.dt_mpc, S.wb, F);
        double acc = 0;
        for (int r = 0; r < 5; ++r) acc += y[L.YCOL(0, r, n)] * F[r].a;
        JG[n * 5 + i] = acc;
        if (i == 0)
          for (int r = 0; r < 5; ++r) c[L.YCOL(0, r, n)] = x[L.Z(0, r, n + 1)] - F[r].v;
      } else {
        const int n = (unit - uPair - uObs) * 32 + ln;
        if (n >= nDyn) continue;
        double z[NZ], F[5];
        for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(0, q, n)];
        if (EU) euler_value(z, S.dt_mpc, S.wb, F);
        else rk4_value(z, S.dt_mpc, S.wb, F);
        for (int r = 0; r < 5; ++r) c[L.YCOL(0, r, n)] = x[L.Z(0, r, n + 1)] - F[r];
      }
    }
  }
  prof_mark(ctx, 0);
  cta_sync(ctx);
  for (int n = ctx.tid; n < N; n += ctx.nt) {
    double z[NZ];
    for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(0, q, n)];
    const double* rf = par.ref + 3 * n;
    double ex = z[0] - rf[0], ey = z[1] - rf[1], ep = z[2] - rf[2];
    if (EU) f_part += z[5] * z[5] + z[6] * z[6];
    else f_part += 100.0 * (ex * ex + ey * ey + ep * ep) + z[5] * z[5] + z[3] * z[3] * z[6] * z[6] + z[4] * z[4];
    for (int o = 0; o < L.P; ++o) f_part += S.rho * x[L.PEL(o, n)];
    if (n == 0) {
      for (int q = 0; q < 5; ++q) c[L.YINIT(0, q)] = z[q] - par.cur[q];
      if (EU) c[L.YINIT(0, 5)] = z[5], c[L.YINIT(0, 6)] = z[6];
    }
    if (EU && n == N - 1 && L.heading[0]) c[L.YTERM(0, 0)] = z[2] - S.heading[0];
    const int qs = EU ? euler_set_at(L, n) : -1;
    double tg[3] = {0, 0, 0};
    if (qs >= 1) {
      const double cs = cos(z[2]), sn = sin(z[2]), fx = z[0] + S.wb * cs, fy = z[1] + S.wb * sn;
      for (int r = 0; r < 4; ++r) {
        const double* tb = S.tube_row(L, 0, qs, 0, r);
        const double* tf = S.tube_row(L, 0, qs, 1, r);
        c[L.YTUBE(0, qs - 1, r)] = tb[2] - tb[0] * z[0] - tb[1] * z[1] - x[L.TS(0, qs - 1, r)];
        c[L.YTUBE(0, qs - 1, 4 + r)] = tf[2] - tf[0] * fx - tf[1] * fy - x[L.TS(0, qs - 1, 4 + r)];
        if (y) {
          const double yb = y[L.YTUBE(0, qs - 1, r)], yf = y[L.YTUBE(0, qs - 1, 4 + r)];
          gl[L.TS(0, qs - 1, r)] = -yb;
          gl[L.TS(0, qs - 1, 4 + r)] = -yf;
          tg[0] -= tb[0] * yb + tf[0] * yf;
          tg[1] -= tb[1] * yb + tf[1] * yf;
          tg[2] -= yf * S.wb * (-tf[0] * sn + tf[1] * cs);
        }
      }
    }
    if (!y) continue;
    double g[NZ] = {200.0 * ex, 200.0 * ey, 200.0 * ep, 2.0 * z[3] * z[6] * z[6], 2.0 * z[4], 2.0 * z[5], 2.0 * z[3] * z[3] * z[6]};
    if (EU) g[0] = tg[0], g[1] = tg[1], g[2] = tg[2], g[3] = 0.0, g[4] = 0.0, g[6] = 2.0 * z[6];
    if (n < N - 1) {
      g[0] -= y[L.YCOL(0, 0, n)], g[1] -= y[L.YCOL(0, 1, n)];  // F_x = x + ..., F_y = y + ...: identity columns
      for (int i = 0; i < 5; ++i) g[2 + i] -= JG[n * 5 + i];
    }
    if (n >= 1)
      for (int r = 0; r < 5; ++r) g[r] += y[L.YCOL(0, r, n - 1)];
    if (n == 0)
      for (int q = 0; q < (EU ? NZ : 5); ++q) g[q] += y[L.YINIT(0, q)];
    if (EU && n == N - 1 && L.heading[0]) g[2] += y[L.YTERM(0, 0)];
    for (int j = 0; j < L.O; ++j) {
      const double* og = OG + (size_t)(n * L.O + j) * 3;
      g[0] += og[0], g[1] += og[1], g[2] += og[2];
    }
    for (int o = 0; o < L.P; ++o) {
      const double* pg = PG + (size_t)o * 6 * L.Mv + n;
      g[0] += pg[0], g[1] += pg[L.Mv], g[2] += pg[2 * L.Mv];
    }
    for (int q = 0; q < NZ; ++q) gl[L.Z(0, q, n)] = g[q];
  }
  *f_out = cta_sum(ctx, f_part);
  *gdt_out = 0.0;
  cta_sync(ctx);
  prof_mark(ctx, 1);
}

// ------------------------------------------------------------------------------------------------
// [LOCAL] + [RICCATI] + [BACKSUB]
// ------------------------------------------------------------------------------------------------
// SM: the work arena RW is shared memory (MPC horizons); false: per-slot global memory (the 271-node state warm start)
template <bool SM>
OBCA_HDN int mpc_kkt_solve_impl(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, double* RW, int* ok_shared) {
  OBCA_ASSUME_STATIC(L, S);
  assume_scratch(W);
  if (SM) OBCA_ASSUME_SHARED(RW);
  else OBCA_ASSUME_GLOBAL(RW);
  const MpcPar par = mpc_par(L, W);
  const int N = L.Mv;
  const bool EU = L.euler != 0;
  const double *x = W.x, *y = W.y;
  if (ctx.tid == 0) *ok_shared = 1;
  cta_sync(ctx);
  prof_mark(ctx, 11);
  // Three independent task families -- pair blocks (the longest chains), obstacle blocks, (node, direction) Taylor tasks of
  // the RK4 map -- are dealt out in units of 32 tasks of one family (no divergence inside a warp), longest first, from a
  // shared counter: the warps stay busy until all families are done instead of waiting at a barrier after each family.
  double* HN = RW;                    // [N][28]
  double* GN = HN + (size_t)N * 28;   // [N][7]
  double* AJ = GN + (size_t)N * 7;    // [N][35]
  double* PP = AJ + (size_t)N * 35;   // [N][25 + 5] cost-to-go
  double* KK = PP + (size_t)N * 30;   // [N][2*5 + 2] gains
  double* DZ = KK + (size_t)N * 12;   // [N][7] step
  double* QD = DZ + (size_t)N * 7;    // [N][15] y-contracted second directional derivatives of the RK4 map
  double* OB = QD + (size_t)N * 15;   // [N][O][9] obstacle Schur complements on the pose (6 sym + 3 grad)
  double* CR_ = OB + (size_t)N * L.O * 9 + 128;  // [N][5] + [5] staged residuals (= CR below, behind the stage scratch)
  const int nPair = L.P * N, nObs = N * L.O, nDir = (N - 1) * 15;
  const int uPair = (nPair + 31) / 32, uObs = (nObs + 31) / 32, uDir = (nDir + 31) / 32;
  int* next_unit = ok_shared + 1;  // Shared::again, unused in MPC mode
  if (ctx.tid == 0) *next_unit = 0;
  cta_sync(ctx);
  for (;;) {
    int unit;
#if defined(__CUDA_ARCH__)
    {
      int u0 = 0;
      if ((ctx.tid & 31) == 0) u0 = atomicAdd(next_unit, 1);
      unit = __shfl_sync(0xffffffffu, u0, 0);
    }
    const int lane0 = ctx.tid & 31, lane1 = lane0 + 1;
#else
    unit = (*next_unit)++;
    const int lane0 = 0, lane1 = 32;
#endif
    if (unit >= uPair + uObs + uDir) break;
    for (int ln = lane0; ln < lane1; ++ln) {
      if (unit < uPair) {
        // pair block: same elimination as the joint problem, the other pose being constant (its Schur block is unused)
        const int it = unit * 32 + ln;
        if (it >= nPair) continue;
        const int o = it / N, n = it % N;
        Pose a, b;
        load_pose(L, x, 0, n, a);
        load_other_pose(L, par, o, n, b);
        pair_block_eliminate(L, S, W, o, n, a, b, ok_shared);
      } else if (unit < uPair + uObs) {
        const int it = (unit - uPair) * 32 + ln;
        if (it >= nObs) continue;
        const int n = it / L.O, j = it % L.O;
        Pose p;
        load_pose(L, x, 0, n, p);
        double H6[6] = {0, 0, 0, 0, 0, 0}, g3[3] = {0, 0, 0};
        obs_block_eliminate(L, S, W, 0, n, j, p, H6, g3, ok_shared);
        double* ob = OB + (size_t)it * 9;
        for (int q = 0; q < 6; ++q) ob[q] = H6[q];
        for (int q = 0; q < 3; ++q) ob[6 + q] = g3[q];
      } else {
        // (node, direction): e_i and e_i + e_j over the nonlinear inputs (psi, v, delta, a, w); x and y enter F linearly
        const int it = (unit - uPair - uObs) * 32 + ln;
        if (it >= nDir) continue;
        const int n = it / 15, k = it % 15;
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= k) ++i;
        const int j = k - i * (i + 1) / 2;
        double z[NZ], d[NZ] = {0, 0, 0, 0, 0, 0, 0};
        for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(0, q, n)];
        d[2 + i] = 1.0, d[2 + j] = 1.0;
        Tay<true> F[5];
        if (EU) tay_euler<true>(z, d, S.dt_mpc, S.wb, F);
        else tay_rk4<true>(z, d, S.dt_mpc, S.wb, F);
        double acc = 0;
        for (int r = 0; r < 5; ++r) acc += y[L.YCOL(0, r, n)] * 2.0 * F[r].b;
        QD[n * 15 + k] = acc;
        if (i == j)
          for (int r = 0; r < 5; ++r) AJ[(size_t)n * 35 + r * 7 + 2 + i] = F[r].a;
      }
    }
  }
  prof_mark(ctx, 2);
  cta_sync(ctx);
  for (int n = ctx.tid; n < N; n += ctx.nt) {
    double z[NZ];
    for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(0, q, n)];
    double H[28], g[NZ];
    for (int q = 0; q < 28; ++q) H[q] = 0;
    for (int q = 0; q < NZ; ++q) H[sym(q, q)] = W.sig[L.Z(0, q, n)], g[q] = W.gphi[L.Z(0, q, n)];
    if (EU) {
      H[sym(5, 5)] += 2.0, H[sym(6, 6)] += 2.0;
      const int qs = euler_set_at(L, n);
      if (qs >= 1) {  // tube set: slack and multiplier eliminated analytically (same algebra as node_assemble, obca_kkt.h)
        const double cs = cos(z[2]), sn = sin(z[2]);
        for (int r = 0; r < 8; ++r) {
          const double* t = S.tube_row(L, 0, qs, r / 4, r % 4);
          const double gr[3] = {t[0], t[1], r < 4 ? 0.0 : S.wb * (-t[0] * sn + t[1] * cs)};
          const double sg = W.sig[L.TS(0, qs - 1, r)];
          const double w = sg * W.c[L.YTUBE(0, qs - 1, r)] + W.gphi[L.TS(0, qs - 1, r)];
          for (int m = 0; m < 3; ++m) {
            g[m] -= gr[m] * w;
            for (int mm = 0; mm <= m; ++mm) H[sym(m, mm)] += sg * gr[m] * gr[mm];
          }
          if (r >= 4) H[sym(2, 2)] += y[L.YTUBE(0, qs - 1, r)] * S.wb * (t[0] * cs + t[1] * sn);
        }
      }
    } else {
      H[sym(0, 0)] += 200.0, H[sym(1, 1)] += 200.0, H[sym(2, 2)] += 200.0;
      H[sym(3, 3)] += 2.0 * z[6] * z[6], H[sym(6, 6)] += 2.0 * z[3] * z[3], H[sym(6, 3)] += 4.0 * z[3] * z[6];
      H[sym(4, 4)] += 2.0, H[sym(5, 5)] += 2.0;
    }
    if (n < N - 1) {
      const double* qd = QD + n * 15;
      for (int i = 0; i < 5; ++i)
        for (int j = 0; j <= i; ++j) {
          const double hij = i == j ? qd[sym(i, i)] : 0.5 * (qd[sym(i, j)] - qd[sym(i, i)] - qd[sym(j, j)]);
          H[sym(2 + i, 2 + j)] -= hij;
        }
      for (int r = 0; r < 5; ++r) AJ[(size_t)n * 35 + r * 7 + 0] = r == 0 ? 1.0 : 0.0, AJ[(size_t)n * 35 + r * 7 + 1] = r == 1 ? 1.0 : 0.0;
    }
    for (int j = 0; j < L.O; ++j) {
      const double* ob = OB + (size_t)(n * L.O + j) * 9;
      for (int q = 0; q < 6; ++q) H[q] += ob[q];
      for (int q = 0; q < 3; ++q) g[q] += ob[6 + q];
    }
    for (int o = 0; o < L.P; ++o) {
      const double* ph = W.PH + (size_t)o * 27 * L.Mv + n;
      for (int r = 0; r < 3; ++r) {
        for (int m = 0; m <= r; ++m) H[sym(r, m)] += ph[sym(r, m) * L.Mv];
        g[r] += ph[(21 + r) * L.Mv];
      }
    }
    for (int q = 0; q < 28; ++q) HN[(size_t)n * 28 + q] = H[q];
    for (int q = 0; q < NZ; ++q) GN[(size_t)n * 7 + q] = g[q];
    if (n < N - 1)
      for (int r = 0; r < 5; ++r) CR_[n * 5 + r] = W.c[L.YCOL(0, r, n)];
    if (n == 0)
      for (int r = 0; r < 5; ++r) CR_[N * 5 + r] = W.c[L.YINIT(0, r)];
  }
  cta_sync(ctx);
  prof_mark(ctx, 3);
  // Riccati recursion (state 5, control 2), serial over the stages:
  // warp 0, lanes over the matrix entries; per stage: PA = P A, Q += A'PA, gains, cost-to-go.
  // State warm start (EU): the inputs of stage 0 are fixed by a_0 = w_0 = 0 (no gain, du_0 = -c), and a final-heading row
  // e'dz_{N-1} = -c_h is handled exactly by a second right-hand side through the same factorisation (TWO): solution 2 answers a unit
  // gradient on psi_{N-1}; the row's multiplier step is dy_h = -(c_h + e'dz1) / (e'dz2) and every quantity is X1 + dy_h X2.
  double* SC = OB + (size_t)N * L.O * 9;  // [128] stage scratch: PA[35], Pr[5], Q[49], qv[7], K[12], Pr2[5], qv2[7]
  double* CR = SC + 128;                  // [N][5] + [5] dynamics / initial-state residuals staged next to the stage data: the serial
                                          // recursion would otherwise wait for a global-memory round trip in every stage
  double* K2 = CR + (size_t)(N + 1) * 5;  // [N][2]   constant gain terms of solution 2
  double* P2 = K2 + (size_t)N * 2;        // [N+1][5] cost-to-go gradient of solution 2
  double* DZ2 = P2 + (size_t)(N + 1) * 5; // [N][7]   step of solution 2
  const bool TWO = EU && L.heading[0];
  if (ctx.tid < 32) {
    double *PA = SC, *Pr = SC + 35, *Q = SC + 40, *qv = SC + 89, *Pr2 = SC + 108, *qv2 = SC + 113;
    for (int n = N - 1; n >= 0; --n) {
      const double* H = HN + (size_t)n * 28;
      const double* g = GN + (size_t)n * 7;
      const double* A = AJ + (size_t)n * 35;
      const double* Pn = PP + (size_t)(n + 1) * 30;  // cost-to-go of the next stage (not read for n = N-1)
      double* K = KK + (size_t)n * 12;
      if (n < N - 1) {
        OBCA_LANES(lane) {
          for (int e = lane; e < (TWO ? 45 : 40); e += 32) {
            if (e < 35) {
              const int r = e / 7, q = e % 7;
              double sacc = 0;
              for (int m = 0; m < 5; ++m) sacc += Pn[r * 5 + m] * A[m * 7 + q];
              PA[e] = sacc;
            } else if (e < 40) {
              const int r = e - 35;
              double cr = 0;
              for (int m = 0; m < 5; ++m) cr += Pn[r * 5 + m] * CR[n * 5 + m];
              Pr[r] = Pn[25 + r] - cr;  // P (-r) + pv
            } else
              Pr2[e - 40] = P2[(size_t)(n + 1) * 5 + (e - 40)];
          }
        }
        OBCA_WARP_SYNC();
      }
      OBCA_LANES(lane) {
        for (int e = lane; e < (TWO ? 63 : 56); e += 32) {
          if (e < 49) {
            const int r = e / 7, q = e % 7;
            double sacc = H[sym(r, q)];
            if (n < N - 1)
              for (int m = 0; m < 5; ++m) sacc += A[m * 7 + r] * PA[m * 7 + q];
            Q[e] = sacc;
          } else if (e < 56) {
            const int r = e - 49;
            double sacc = g[r];
            if (n < N - 1)
              for (int m = 0; m < 5; ++m) sacc += A[m * 7 + r] * Pr[m];
            qv[r] = sacc;
          } else {
            const int r = e - 56;
            double sacc = (n == N - 1 && r == 2) ? 1.0 : 0.0;
            if (n < N - 1)
              for (int m = 0; m < 5; ++m) sacc += A[m * 7 + r] * Pr2[m];
            qv2[r] = sacc;
          }
        }
      }
      OBCA_WARP_SYNC();
      // eliminate the control (rows/cols 5,6): F = Q_uu must be positive definite
      double f00 = Q[5 * 7 + 5], f01 = Q[5 * 7 + 6], f11 = Q[6 * 7 + 6], det = f00 * f11 - f01 * f01;
      if (!(f00 > 0) || !(det > 1e-14 * f00 * f11)) {
        *ok_shared = 0;
#ifdef OBCA_HOST_EMU
        if (getenv("OBCA_TRACE")) printf("       stage %d: F not positive definite (%.3e %.3e %.3e)\n", n, f00, f01, f11);
#endif
        f00 = f11 = 1.0, f01 = 0.0, det = 1.0;
      }
      const double i00 = f11 / det, i01 = -f01 / det, i11 = f00 / det;
      const bool fixed_u = EU && n == 0;
      OBCA_LANES(lane) {
        if (lane < 12) {
          const int q = lane % 6, row = lane / 6;  // q = 5: the constant term
          const double b5 = q < 5 ? Q[5 * 7 + q] : qv[5], b6 = q < 5 ? Q[6 * 7 + q] : qv[6];
          double kv = row == 0 ? -(i00 * b5 + i01 * b6) : -(i01 * b5 + i11 * b6);
          if (fixed_u) kv = q < 5 ? 0.0 : -W.c[L.YINIT(0, 5 + row)];
          K[q < 5 ? row * 5 + q : 10 + row] = kv;
        } else if (TWO && lane < 14) {
          const int row = lane - 12;
          const double kv = row == 0 ? -(i00 * qv2[5] + i01 * qv2[6]) : -(i01 * qv2[5] + i11 * qv2[6]);
          K2[(size_t)n * 2 + row] = fixed_u ? 0.0 : kv;
        }
      }
      OBCA_WARP_SYNC();
      OBCA_LANES(lane) {
        for (int e = lane; e < (TWO ? 35 : 30); e += 32) {
          double* Po = PP + (size_t)n * 30;
          if (e < 25) {
            const int r = e / 5, q = e % 5;
            Po[e] = Q[r * 7 + q] + Q[r * 7 + 5] * K[q] + Q[r * 7 + 6] * K[5 + q];
          } else if (e < 30) {
            const int r = e - 25;
            Po[e] = qv[r] + Q[r * 7 + 5] * K[10] + Q[r * 7 + 6] * K[11];
          } else {
            const int r = e - 30;
            P2[(size_t)n * 5 + r] = qv2[r] + Q[r * 7 + 5] * K2[(size_t)n * 2] + Q[r * 7 + 6] * K2[(size_t)n * 2 + 1];
          }
        }
      }
      OBCA_WARP_SYNC();
    }
    // forward pass
    OBCA_LANES(lane) {
      if (lane < 5) DZ[lane] = -CR[N * 5 + lane];
      else if (TWO && lane < 10) DZ2[lane - 5] = 0.0;
    }
    OBCA_WARP_SYNC();
    for (int n = 0; n < N; ++n) {
      const double* K = KK + (size_t)n * 12;
      double* w = DZ + (size_t)n * 7;
      double* w2 = DZ2 + (size_t)n * 7;
      OBCA_LANES(lane) {
        if (lane < 2) {
          double du = K[10 + lane];
          for (int q = 0; q < 5; ++q) du += K[lane * 5 + q] * w[q];
          w[5 + lane] = du;
        } else if (TWO && lane < 4) {
          const int row = lane - 2;
          double du = K2[(size_t)n * 2 + row];
          for (int q = 0; q < 5; ++q) du += K[row * 5 + q] * w2[q];
          w2[5 + row] = du;
        }
      }
      OBCA_WARP_SYNC();
      if (n < N - 1) {
        const double* A = AJ + (size_t)n * 35;
        OBCA_LANES(lane) {
          if (lane < 5) {
            double sacc = -CR[n * 5 + lane];
            for (int q = 0; q < 7; ++q) sacc += A[lane * 7 + q] * w[q];
            w[7 + lane] = sacc;  // dz of stage n+1
          } else if (TWO && lane < 10) {
            const int r = lane - 5;
            double sacc = 0.0;
            for (int q = 0; q < 7; ++q) sacc += A[r * 7 + q] * w2[q];
            w2[7 + r] = sacc;
          }
        }
        OBCA_WARP_SYNC();
      }
    }
    if (EU) {
      // combine the two solutions and recover the multipliers of the a_0 = w_0 = 0 rows from the stationarity of u_0
      // (Q, qv, qv2 still hold stage 0): dy_u = -(Q_u. [dz_0; du_0] + q_u)
      double dyh = 0.0;
      if (TWO) {
        const double e1 = DZ[(size_t)(N - 1) * 7 + 2], e2 = DZ2[(size_t)(N - 1) * 7 + 2];
#ifdef OBCA_HOST_EMU
        if (getenv("OBCA_TRACE")) printf("       heading row: e1 %.3e e2 %.3e c %.3e\n", e1, e2, W.c[L.YTERM(0, 0)]);
#endif
        // e'dz2 = -(e' Hred^-1 e) <= 0; it vanishes where the linearised dynamics cannot turn the vehicle (v = 0 and delta = 0 along the
        // whole guess): the pivot is then bounded away from zero, the role of IPOPT's delta_c perturbation of a singular KKT matrix
        if (!(e2 <= 0.0)) *ok_shared = 0;
        else dyh = -(W.c[L.YTERM(0, 0)] + e1) / fmin(e2, -1e-8);
        OBCA_LANES(lane) {
          for (int e = lane; e < N * 7; e += 32) DZ[e] += dyh * DZ2[e];
        }
        OBCA_LANES(lane) {
          for (int e = lane; e < N * 5; e += 32) PP[(size_t)(e / 5) * 30 + 25 + e % 5] += dyh * P2[e];
        }
        OBCA_WARP_SYNC();
      }
      OBCA_LANES(lane) {
        if (lane < 2) {
          double sacc = qv[5 + lane] + (TWO ? dyh * qv2[5 + lane] : 0.0);
          for (int q = 0; q < 7; ++q) sacc += Q[(5 + lane) * 7 + q] * DZ[q];
          W.dy[L.YINIT(0, 5 + lane)] = -sacc;
        } else if (lane == 2 && TWO)
          W.dy[L.YTERM(0, 0)] = dyh;
      }
      OBCA_WARP_SYNC();
    }
  }
  cta_sync(ctx);
  prof_mark(ctx, 6);
  if (!*ok_shared) return 0;
  // primal step, multipliers (costates), local blocks
  for (int n = ctx.tid; n < N; n += ctx.nt) {
    const double* w = DZ + (size_t)n * 7;
    for (int q = 0; q < NZ; ++q) W.dx[L.Z(0, q, n)] = w[q];
    // multiplier of the row that defines z_n: dy = -(P_n dz_n + p_n)
    const double* P = PP + (size_t)n * 30;
    for (int r = 0; r < 5; ++r) {
      double s = P[25 + r];
      for (int m = 0; m < 5; ++m) s += P[r * 5 + m] * w[m];
      if (n == 0) W.dy[L.YINIT(0, r)] = -s;
      else W.dy[L.YCOL(0, r, n - 1)] = -s;
    }
    if (EU) {
      const int qs = euler_set_at(L, n);
      if (qs >= 1) {
        const double psi = x[L.Z(0, 2, n)], cs = cos(psi), sn = sin(psi);
        for (int r = 0; r < 8; ++r) {
          const double* t = S.tube_row(L, 0, qs, r / 4, r % 4);
          const double gr[3] = {t[0], t[1], r < 4 ? 0.0 : S.wb * (-t[0] * sn + t[1] * cs)};
          const double dts = W.c[L.YTUBE(0, qs - 1, r)] - (gr[0] * w[0] + gr[1] * w[1] + gr[2] * w[2]);
          W.dx[L.TS(0, qs - 1, r)] = dts;
          W.dy[L.YTUBE(0, qs - 1, r)] = W.sig[L.TS(0, qs - 1, r)] * dts + W.gphi[L.TS(0, qs - 1, r)];
        }
      }
    }
  }
  // local blocks: one task per (node, obstacle) and (other, node)
  for (int it = ctx.tid; it < N * (L.O + L.P); it += ctx.nt) {
    if (it < N * L.O) {
      const int n = it / L.O, j = it % L.O;
      const double* w = DZ + (size_t)n * 7;
      double dp[3] = {w[0], w[1], w[2]};
      obs_block_backsub(L, W, 0, n, j, dp);
    } else {
      const int e = it - N * L.O, o = e / N, n = e % N;
      const double* w = DZ + (size_t)n * 7;
      double dp6[6] = {w[0], w[1], w[2], 0.0, 0.0, 0.0};
      pair_block_backsub(L, W, o, n, dp6);
    }
  }
  cta_sync(ctx);
  prof_mark(ctx, 10);
  return 1;
}

OBCA_HDN int mpc_kkt_solve(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, double* RW, int* ok_shared) {
  if (W.ricg) return mpc_kkt_solve_impl<false>(ctx, L, S, W, W.ricg, ok_shared);
  return mpc_kkt_solve_impl<true>(ctx, L, S, W, RW, ok_shared);
}

inline size_t mpc_work_doubles(const Lay& L) { return (size_t)(L.Mv + 1) * (28 + 7 + 35 + 30 + 12 + 7 + 15 + 5 + 9 * L.O) + 128 + (L.euler ? (size_t)(L.Mv + 1) * 14 : 0); }

}  // namespace obca
