// obca_refine.h -- matrix-free application of the primal-dual Newton matrix and iterative refinement of the structured
// solve (collocation mode).
//
//     r1 = (W + Sigma + dw I) dx + J' dy         (x-layout; W.sig already holds Sigma + dw)
//     r2 =  J dx - dc dy                         (y-layout; dc = DELTA_C_LOCAL on obstacle / pair rows, 0 elsewhere)
//
// W is the Lagrangian Hessian with the same clipped norm-row multipliers as the eliminations in obca_kkt.h, but it is
// written down block by block from the constraint functions and shares no code with them: kkt_apply is both the
// residual evaluation of the refinement loop (IPOPT's iterative refinement on the full augmented system) and an
// independent check of the structured solve (tests: K [dx; dy] against the oracle's sparse matrix).
#pragma once

namespace obca {

// ---- pair phase: local rows / variables of every (pair, node) block; pose contributions -> PG[p][n][6]
OBCA_HDN void apply_pairs(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, const double* dx, const double* dy, double* r1, double* r2) {
  OBCA_ASSUME_STATIC(L, S);
  assume_scratch(W);
  OBCA_ASSUME_GLOBAL(dx), OBCA_ASSUME_GLOBAL(dy), OBCA_ASSUME_GLOBAL(r1), OBCA_ASSUME_GLOBAL(r2);
  const double *x = W.x, *y = W.y;
  for (int it = ctx.tid; it < L.nPairNodes; it += ctx.nt) {
    int p = 0, n = it;
    while (n >= L.Mp[p]) n -= L.Mp[p], ++p;
    const int ia = L.pa[p], ib = L.pb[p];
    Pose a, b;
    load_pose(L, x, ia, n, a);
    load_pose(L, x, ib, n, b);
    PairBlk B;
    load_pair(L, x, p, n, B);
    pair_residual(S, a, b, B);
    const double yd = y[L.YPAIR(p, 0, n)], ye1[2] = {y[L.YPAIR(p, 1, n)], y[L.YPAIR(p, 2, n)]};
    const double ye2[2] = {y[L.YPAIR(p, 3, n)], y[L.YPAIR(p, 4, n)]}, yn = y[L.YPAIR(p, 5, n)];
    const double ynm = yn < 0 ? yn : 0.0;
    double dl[4], dm[4];
    for (int r = 0; r < 4; ++r) dl[r] = dx[L.PL(p, r, n)], dm[r] = dx[L.PM(p, r, n)];
    const double ds[2] = {dx[L.PS(p, 0, n)], dx[L.PS(p, 1, n)]};
    const double dsd = dx[L.PSD(p, n)], dsn = dx[L.PSN(p, n)], del = dx[L.PEL(p, n)];
    const double dpa[3] = {dx[L.Z(ia, 0, n)], dx[L.Z(ia, 1, n)], dx[L.Z(ia, 2, n)]};
    const double dpb[3] = {dx[L.Z(ib, 0, n)], dx[L.Z(ib, 1, n)], dx[L.Z(ib, 2, n)]};
    double dyv[6];
    for (int r = 0; r < 6; ++r) dyv[r] = dy[L.YPAIR(p, r, n)];
    // R' u (R' = dR/dpsi) and t . R' u
    const double dRua[2] = {-a.s * B.ua[0] - a.c * B.ua[1], a.c * B.ua[0] - a.s * B.ua[1]};
    const double dRub[2] = {-b.s * B.ub[0] - b.c * B.ub[1], b.c * B.ub[0] - b.s * B.ub[1]};
    const double tdRa = a.x * dRua[0] + a.y * dRua[1], tdRb = b.x * dRub[0] + b.y * dRub[1];
    // rows
    double jd = -B.Rua[0] * dpa[0] - B.Rua[1] * dpa[1] - tdRa * dpa[2] - B.Rub[0] * dpb[0] - B.Rub[1] * dpb[1] - tdRb * dpb[2] - dsd + del;
    double je1[2] = {dRua[0] * dpa[2] + ds[0], dRua[1] * dpa[2] + ds[1]};
    double je2[2] = {dRub[0] * dpb[2] - ds[0], dRub[1] * dpb[2] - ds[1]};
    const double jn = -2.0 * (B.s[0] * ds[0] + B.s[1] * ds[1]) - dsn;
    double ga[3] = {-B.Rua[0] * dyv[0], -B.Rua[1] * dyv[0], -tdRa * dyv[0] + dRua[0] * dyv[1] + dRua[1] * dyv[2]};
    double gb[3] = {-B.Rub[0] * dyv[0], -B.Rub[1] * dyv[0], -tdRb * dyv[0] + dRub[0] * dyv[3] + dRub[1] * dyv[4]};
    for (int r = 0; r < 4; ++r) {
      const double g0 = S.G[r][0], g1 = S.G[r][1];
      // R G_r' and R' G_r' for both poses
      const double ea = a.c * g0 - a.s * g1, fa = a.s * g0 + a.c * g1, eb = b.c * g0 - b.s * g1, fb = b.s * g0 + b.c * g1;
      const double dax = -a.s * g0 - a.c * g1, day = a.c * g0 - a.s * g1, dbx = -b.s * g0 - b.c * g1, dby = b.c * g0 - b.s * g1;
      jd -= B.ba[r] * dl[r] + B.bb[r] * dm[r];
      je1[0] += ea * dl[r], je1[1] += fa * dl[r];
      je2[0] += eb * dm[r], je2[1] += fb * dm[r];
      // Hessian couplings (lam_r, pose a) and (mu_r, pose b)
      const double ca0 = -yd * ea, ca1 = -yd * fa, ca2 = -yd * (dax * a.x + day * a.y) + ye1[0] * dax + ye1[1] * day;
      const double cb0 = -yd * eb, cb1 = -yd * fb, cb2 = -yd * (dbx * b.x + dby * b.y) + ye2[0] * dbx + ye2[1] * dby;
      r1[L.PL(p, r, n)] = W.sig[L.PL(p, r, n)] * dl[r] + ca0 * dpa[0] + ca1 * dpa[1] + ca2 * dpa[2] - B.ba[r] * dyv[0] + ea * dyv[1] + fa * dyv[2];
      r1[L.PM(p, r, n)] = W.sig[L.PM(p, r, n)] * dm[r] + cb0 * dpb[0] + cb1 * dpb[1] + cb2 * dpb[2] - B.bb[r] * dyv[0] + eb * dyv[3] + fb * dyv[4];
      ga[0] += ca0 * dl[r], ga[1] += ca1 * dl[r], ga[2] += ca2 * dl[r];
      gb[0] += cb0 * dm[r], gb[1] += cb1 * dm[r], gb[2] += cb2 * dm[r];
    }
    // pose-pose curvature
    ga[0] += -yd * dRua[0] * dpa[2], ga[1] += -yd * dRua[1] * dpa[2];
    ga[2] += -yd * (dRua[0] * dpa[0] + dRua[1] * dpa[1]) + (yd * (a.x * B.Rua[0] + a.y * B.Rua[1]) - (ye1[0] * B.Rua[0] + ye1[1] * B.Rua[1])) * dpa[2];
    gb[0] += -yd * dRub[0] * dpb[2], gb[1] += -yd * dRub[1] * dpb[2];
    gb[2] += -yd * (dRub[0] * dpb[0] + dRub[1] * dpb[1]) + (yd * (b.x * B.Rub[0] + b.y * B.Rub[1]) - (ye2[0] * B.Rub[0] + ye2[1] * B.Rub[1])) * dpb[2];
    r1[L.PS(p, 0, n)] = (W.sig[L.PS(p, 0, n)] - 2.0 * ynm) * ds[0] + dyv[1] - dyv[3] - 2.0 * B.s[0] * dyv[5];
    r1[L.PS(p, 1, n)] = (W.sig[L.PS(p, 1, n)] - 2.0 * ynm) * ds[1] + dyv[2] - dyv[4] - 2.0 * B.s[1] * dyv[5];
    r1[L.PSD(p, n)] = W.sig[L.PSD(p, n)] * dsd - dyv[0];
    r1[L.PSN(p, n)] = W.sig[L.PSN(p, n)] * dsn - dyv[5];
    r1[L.PEL(p, n)] = W.sig[L.PEL(p, n)] * del + dyv[0];
    r2[L.YPAIR(p, 0, n)] = jd - DELTA_C_LOCAL * dyv[0];
    r2[L.YPAIR(p, 1, n)] = je1[0] - DELTA_C_LOCAL * dyv[1];
    r2[L.YPAIR(p, 2, n)] = je1[1] - DELTA_C_LOCAL * dyv[2];
    r2[L.YPAIR(p, 3, n)] = je2[0] - DELTA_C_LOCAL * dyv[3];
    r2[L.YPAIR(p, 4, n)] = je2[1] - DELTA_C_LOCAL * dyv[4];
    r2[L.YPAIR(p, 5, n)] = jn - DELTA_C_LOCAL * dyv[5];
    double* g = W.PG + (size_t)p * 6 * L.Mv + n;
    const int gs = L.Mv;
    g[0] = ga[0], g[gs] = ga[1], g[2 * gs] = ga[2], g[3 * gs] = gb[0], g[4 * gs] = gb[1], g[5 * gs] = gb[2];
  }
}

// ---- node phase: states, obstacle / tube blocks, collocation / continuity / boundary rows, the dt row
OBCA_HDN void apply_nodes(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, const double* dx, const double* dy, double* r1, double* r2) {
  OBCA_ASSUME_STATIC(L, S);
  assume_scratch(W);
  OBCA_ASSUME_GLOBAL(dx), OBCA_ASSUME_GLOBAL(dy), OBCA_ASSUME_GLOBAL(r1), OBCA_ASSUME_GLOBAL(r2);
  const double *x = W.x, *y = W.y;
  const double dt = x[L.oDT], idt = 1.0 / dt, ddt = dx[L.oDT];
  double rdt = 0;  // dt row of r1 (partial sums)
  for (int it = ctx.tid; it < L.V * L.Mv; it += ctx.nt) {
    const int a = it / L.Mv, n = it % L.Mv;
    if (n >= L.M[a]) continue;
    const int i = n / NK, k = n % NK, n0 = i * NK;
    double z[NZ], dz[NZ], g[NZ];
    for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(a, q, n)], dz[q] = dx[L.Z(a, q, n)], g[q] = W.sig[L.Z(a, q, n)] * dz[q];
    const double cs = cos(z[2]), sn = sin(z[2]), v = z[3], de = z[4], ua = z[5], uw = z[6];
    const double tde = tan(de), sec2 = 1.0 + tde * tde;
    const double bk = S.cB[k], bdt = bk * dt;
    // running cost  bk dt (ua^2 + v^2 uw^2 + de^2)
    g[3] += bdt * (2.0 * uw * uw * dz[3] + 4.0 * v * uw * dz[6]);
    g[6] += bdt * (4.0 * v * uw * dz[3] + 2.0 * v * v * dz[6]);
    g[4] += bdt * 2.0 * dz[4];
    g[5] += bdt * 2.0 * dz[5];
    {
      const double h3 = bk * 2.0 * v * uw * uw, h4 = bk * 2.0 * de, h5 = bk * 2.0 * ua, h6 = bk * 2.0 * v * v * uw;
      g[3] += h3 * ddt, g[4] += h4 * ddt, g[5] += h5 * ddt, g[6] += h6 * ddt;
      rdt += h3 * dz[3] + h4 * dz[4] + h5 * dz[5] + h6 * dz[6];
    }
    // collocation rows (q, n):  sum_j cA[j][k] z_{j,q} / dt - f_q(z_n)
    double yc[5], dyc[5];
    for (int q = 0; q < 5; ++q) {
      yc[q] = y[L.YCOL(a, q, n)], dyc[q] = dy[L.YCOL(a, q, n)];
      double pl = 0, dpl = 0, sy = 0, sdy = 0;
      for (int j = 0; j < NK; ++j) {
        pl += S.cA[j][k] * x[L.Z(a, q, n0 + j)], dpl += S.cA[j][k] * dx[L.Z(a, q, n0 + j)];
        sy += S.cA[k][j] * y[L.YCOL(a, q, n0 + j)], sdy += S.cA[k][j] * dy[L.YCOL(a, q, n0 + j)];
      }
      const double poly = pl * idt;
      double jrow = dpl * idt - poly * idt * ddt;
      if (q == 0) jrow -= -v * sn * dz[2] + cs * dz[3];
      else if (q == 1) jrow -= v * cs * dz[2] + sn * dz[3];
      else if (q == 2) jrow -= tde / S.wb * dz[3] + v * sec2 / S.wb * dz[4];
      else if (q == 3) jrow -= dz[5];
      else jrow -= dz[6];
      r2[L.YCOL(a, q, n)] = jrow;
      g[q] += sdy * idt - sy * idt * idt * ddt;       // J'dy and the (z, dt) curvature of the rows of this interval
      rdt += -dyc[q] * poly * idt - sy * idt * idt * dz[q] + yc[q] * 2.0 * poly * idt * idt * ddt;
    }
    g[2] -= dyc[0] * (-v * sn) + dyc[1] * (v * cs);
    g[3] -= dyc[0] * cs + dyc[1] * sn + dyc[2] * tde / S.wb;
    g[4] -= dyc[2] * v * sec2 / S.wb;
    g[5] -= dyc[3];
    g[6] -= dyc[4];
    // -y . d2f
    g[2] += (yc[0] * v * cs + yc[1] * v * sn) * dz[2] + (yc[0] * sn - yc[1] * cs) * dz[3];
    g[3] += (yc[0] * sn - yc[1] * cs) * dz[2] - yc[2] * sec2 / S.wb * dz[4];
    g[4] += -yc[2] * sec2 / S.wb * dz[3] - yc[2] * 2.0 * v * sec2 * tde / S.wb * dz[4];
    // continuity, initial and terminal rows
    if (k == 0 && i >= 1)
      for (int q = 0; q < NZ; ++q) {
        r2[L.YCONT(a, q, i)] = dx[L.Z(a, q, n - 1)] - dz[q];
        g[q] -= dy[L.YCONT(a, q, i)];
      }
    if (k == NK - 1 && i < L.N[a] - 1)
      for (int q = 0; q < NZ; ++q) g[q] += dy[L.YCONT(a, q, i + 1)];
    if (n == 0)
      for (int q = 0; q < NZ; ++q) r2[L.YINIT(a, q)] = dz[q], g[q] += dy[L.YINIT(a, q)];
    if (n == L.M[a] - 1) {
      r2[L.YTERM(a, 0)] = L.heading[a] ? dz[2] : 0.0;
      if (L.heading[a]) g[2] += dy[L.YTERM(a, 0)];
      for (int q = 3; q < NZ; ++q) r2[L.YTERM(a, q - 2)] = dz[q], g[q] += dy[L.YTERM(a, q - 2)];
    }
    // obstacle blocks
    for (int j = 0; j < L.O; ++j) {
      const double(*A)[2] = S.obsA[j];
      double lam[4], dl[4], dm[4], u0 = 0, u1 = 0, du0 = 0, du1 = 0;
      for (int r = 0; r < 4; ++r) {
        lam[r] = x[L.LAM(a, j, r, n)], dl[r] = dx[L.LAM(a, j, r, n)], dm[r] = dx[L.MU(a, j, r, n)];
        u0 += A[r][0] * lam[r], u1 += A[r][1] * lam[r];
        du0 += A[r][0] * dl[r], du1 += A[r][1] * dl[r];
      }
      const double dsd = dx[L.SD(a, j, n)], del = dx[L.EL(a, j, n)];
      const double y1 = y[L.YOBS(a, j, 0, n)], y2[2] = {y[L.YOBS(a, j, 1, n)], y[L.YOBS(a, j, 2, n)]}, y3 = y[L.YOBS(a, j, 3, n)];
      const double y3p = y3 > 0 ? y3 : 0.0;
      const double d1 = dy[L.YOBS(a, j, 0, n)], d2[2] = {dy[L.YOBS(a, j, 1, n)], dy[L.YOBS(a, j, 2, n)]}, d3 = dy[L.YOBS(a, j, 3, n)];
      // R' = [[c, s], [-s, c]] (transpose of the rotation);  dR'/dpsi u ;  (dR/dpsi) y2
      const double dRtu[2] = {-sn * u0 + cs * u1, -cs * u0 - sn * u1};
      const double dRy[2] = {-sn * y2[0] - cs * y2[1], cs * y2[0] - sn * y2[1]};
      double j1 = u0 * dz[0] + u1 * dz[1] - dsd + del;
      double j2[2] = {cs * du0 + sn * du1 + dRtu[0] * dz[2], -sn * du0 + cs * du1 + dRtu[1] * dz[2]};
      const double j3 = 2.0 * (u0 * du0 + u1 * du1);
      double gpsi = 0;
      for (int r = 0; r < 4; ++r) {
        const double atb = A[r][0] * z[0] + A[r][1] * z[1] - S.obsb[j][r];
        const double jl1 = cs * A[r][0] + sn * A[r][1], jl2 = -sn * A[r][0] + cs * A[r][1], jl3 = 2.0 * (A[r][0] * u0 + A[r][1] * u1);
        const double cl2 = A[r][0] * dRy[0] + A[r][1] * dRy[1];
        j1 += atb * dl[r] - S.g[r] * dm[r];
        j2[0] += S.G[r][0] * dm[r], j2[1] += S.G[r][1] * dm[r];
        r1[L.LAM(a, j, r, n)] = W.sig[L.LAM(a, j, r, n)] * dl[r] + 2.0 * y3p * (A[r][0] * du0 + A[r][1] * du1) + y1 * (A[r][0] * dz[0] + A[r][1] * dz[1]) +
                                cl2 * dz[2] + atb * d1 + jl1 * d2[0] + jl2 * d2[1] + jl3 * d3;
        r1[L.MU(a, j, r, n)] = W.sig[L.MU(a, j, r, n)] * dm[r] - S.g[r] * d1 + S.G[r][0] * d2[0] + S.G[r][1] * d2[1];
        gpsi += cl2 * dl[r];
      }
      r1[L.SD(a, j, n)] = W.sig[L.SD(a, j, n)] * dsd - d1;
      r1[L.EL(a, j, n)] = W.sig[L.EL(a, j, n)] * del + d1;
      r2[L.YOBS(a, j, 0, n)] = j1 - DELTA_C_LOCAL * d1;
      r2[L.YOBS(a, j, 1, n)] = j2[0] - DELTA_C_LOCAL * d2[0];
      r2[L.YOBS(a, j, 2, n)] = j2[1] - DELTA_C_LOCAL * d2[1];
      r2[L.YOBS(a, j, 3, n)] = j3 - DELTA_C_LOCAL * d3;
      g[0] += y1 * du0 + u0 * d1;
      g[1] += y1 * du1 + u1 * d1;
      g[2] += gpsi + dRtu[0] * d2[0] + dRtu[1] * d2[1] - (y2[0] * (cs * u0 + sn * u1) + y2[1] * (-sn * u0 + cs * u1)) * dz[2];
    }
    // tube rows
    const int qs = tube_set_at(L, a, n);
    if (qs >= 1) {
      for (int r = 0; r < 8; ++r) {
        const double* t = S.tube_row(L, a, qs, r / 4, r % 4);
        const double gr2 = r < 4 ? 0.0 : S.wb * (-t[0] * sn + t[1] * cs);
        const double dts = dx[L.TS(a, qs - 1, r)], dyr = dy[L.YTUBE(a, qs - 1, r)];
        r2[L.YTUBE(a, qs - 1, r)] = -(t[0] * dz[0] + t[1] * dz[1] + gr2 * dz[2]) - dts;
        r1[L.TS(a, qs - 1, r)] = W.sig[L.TS(a, qs - 1, r)] * dts - dyr;
        g[0] -= t[0] * dyr, g[1] -= t[1] * dyr, g[2] -= gr2 * dyr;
        if (r >= 4) g[2] += y[L.YTUBE(a, qs - 1, r)] * S.wb * (t[0] * cs + t[1] * sn) * dz[2];
      }
    }
    for (int pp = 0; pp < L.P; ++pp) {
      if (n >= L.Mp[pp]) continue;
      const double* pg = W.PG + (size_t)pp * 6 * L.Mv + n;
      const int gs = L.Mv;
      if (L.pa[pp] == a) g[0] += pg[0], g[1] += pg[gs], g[2] += pg[2 * gs];
      if (L.pb[pp] == a) g[0] += pg[3 * gs], g[1] += pg[4 * gs], g[2] += pg[5 * gs];
    }
    for (int q = 0; q < NZ; ++q) r1[L.Z(a, q, n)] = g[q];
  }
  rdt = cta_sum(ctx, rdt);
  if (ctx.tid == 0) {
    double h = W.sig[L.oDT];
    for (int a = 0; a < L.V; ++a) h += 2.0 * L.N[a] * L.N[a];
    r1[L.oDT] = rdt + h * ddt;
  }
}

// rows / variables of padding nodes and unused tube sets carry no equation: their outputs are zero
OBCA_HDN void apply_clear(const Ctx& ctx, const Lay& L, double* r1, double* r2) {
  for (int q = ctx.tid; q < L.nx; q += ctx.nt) r1[q] = 0;
  for (int q = ctx.tid; q < L.ny; q += ctx.nt) r2[q] = 0;
}

OBCA_HDN void kkt_apply(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, const double* dx, const double* dy, double* r1, double* r2) {
  apply_clear(ctx, L, r1, r2);
  cta_sync(ctx);
  apply_pairs(ctx, L, S, W, dx, dy, r1, r2);
  cta_sync(ctx);
  apply_nodes(ctx, L, S, W, dx, dy, r1, r2);
  cta_sync(ctx);
}

// Iterative refinement of (W.dx, W.dy) = -K^-1 (W.gphi, W.c): residual rho = (gphi, c) + K (dx, dy); when its max-norm
// exceeds `ratio` x the max-norm of the right-hand side, the structured solve is repeated for -rho and the correction is
// added (the factorisation is recomputed: the solve is fused with it).  Work vectors: W.dzL / W.dzU (x-layout),
// W.ct / W.ry (y-layout).  Returns the number of correction solves; *res_out = final residual ratio.
OBCA_HDN int kkt_refine(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, double* RW, int* ok_shared, int max_steps, double ratio, double* res_out) {
  int steps = 0;
  double rr = 0;
  for (;;) {
    kkt_apply(ctx, L, S, W, W.dx, W.dy, W.dzL, W.ct);
    double rmax = 0, bmax = 0;
    for (int q = ctx.tid; q < L.nx; q += ctx.nt) {
      const double r = W.dzL[q] + W.gphi[q];
      W.dzL[q] = r;
      rmax = fmax(rmax, fabs(r)), bmax = fmax(bmax, fabs(W.gphi[q]));
    }
    for (int q = ctx.tid; q < L.ny; q += ctx.nt) {
      const double r = W.ct[q] + W.c[q];
      W.ct[q] = r;
      rmax = fmax(rmax, fabs(r)), bmax = fmax(bmax, fabs(W.c[q]));
    }
    rmax = cta_max(ctx, rmax), bmax = cta_max(ctx, bmax);
    rr = rmax / fmax(bmax, 1e-300);
    if (!(rr > ratio) || steps >= max_steps) break;
    Scratch W2 = W;
    W2.gphi = W.dzL, W2.c = W.ct, W2.dx = W.dzU, W2.dy = W.ry;
    if (!kkt_solve(ctx, L, S, W2, RW, ok_shared)) break;
    for (int q = ctx.tid; q < L.nx; q += ctx.nt) W.dx[q] += W.dzU[q];
    for (int q = ctx.tid; q < L.ny; q += ctx.nt) W.dy[q] += W.ry[q];
    cta_sync(ctx);
    ++steps;
  }
  *res_out = rr;
  return steps;
}

}  // namespace obca
