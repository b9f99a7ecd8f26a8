// obca_kkt.h -- structured solve of the primal-dual Newton system
//
//     [ W + Sigma + dw I   J' ] [dx]     [ gphi ]
//     [ J              -dc I ] [dy] = - [  c   ]
//
// (dc = DELTA_C_LOCAL on obstacle / pair rows only, 0 elsewhere) by
//   [LOCAL]   exact block elimination of the obstacle, pair and tube variables onto the vehicle poses,
//   [NULLSP]  a Householder null-space parametrisation of every (vehicle, interval) collocation block,
//   [RICCATI] a Riccati recursion over the intervals with state (xi_1..xi_V, dt) and control (p_1..p_V).
// The reduced Hessian is positive definite (KKT inertia (n, m, 0), IPOPT's acceptance test) iff every
// Riccati block F_i admits a Cholesky factor and the final dt pivot is positive; otherwise *ok = 0 and
// the caller raises delta_w.
#pragma once

namespace obca {

struct KktAux {
  double* RW;  // Riccati work area
};


// ------------------------------------------------------------------------------------------------
// [LOCAL] pair blocks: one thread per (pair, node)
// unknown order: lam 0-3, mu 4-7 | yd 8, ye1 9-10, ye2 11-12, yn 13 | s 14-15
// The slacks sd, sn and the elastic variable el enter one row each with coefficient -1 / +1 and a
// diagonal Hessian, so they are eliminated analytically into the (yd, yd) and (yn, yn) pivots.
// ------------------------------------------------------------------------------------------------
OBCA_HD void pair_block_eliminate(const Lay& L, const Stat& S, const Scratch& W, int p, int n, const Pose& a, const Pose& b, int* ok) {
  const double *x = W.x, *y = W.y;
  {
    PairBlk B;
    load_pair(L, x, p, n, B);
    pair_residual(S, a, b, B);
    double yd = y[L.YPAIR(p, 0, n)], ye1[2] = {y[L.YPAIR(p, 1, n)], y[L.YPAIR(p, 2, n)]};
    double ye2[2] = {y[L.YPAIR(p, 3, n)], y[L.YPAIR(p, 4, n)]}, yn = y[L.YPAIR(p, 5, n)];
    double ynm = yn < 0 ? yn : 0.0;  // local convexification: exact at KKT points (yn = -z_sn <= 0)
    // Block elimination in registers: lam, mu (diagonal Hessians) -> 5 x 5 Schur complement on (yd, ye1, ye2)
    // (the yn row only has its own pivot), then the 2 x 2 system of s.  7 right-hand sides (6 pose couplings + residual).
    const double isd = rcp_pos(W.sig[L.PSD(p, n)]), iel = rcp_pos(W.sig[L.PEL(p, n)]), isn = rcp_pos(W.sig[L.PSN(p, n)]);
    const double dn = DELTA_C_LOCAL + isn, idn = rcp_pos(dn);
    const double hs0 = W.sig[L.PS(p, 0, n)] - 2.0 * ynm, hs1 = W.sig[L.PS(p, 1, n)] - 2.0 * ynm;
    double sl[4], smu[4], ea[4], fa[4], eb[4], fb[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      sl[r] = rcp_pos(W.sig[L.PL(p, r, n)]);
      smu[r] = rcp_pos(W.sig[L.PM(p, r, n)]);
      ea[r] = a.c * S.G[r][0] - a.s * S.G[r][1];
      fa[r] = a.s * S.G[r][0] + a.c * S.G[r][1];
      eb[r] = b.c * S.G[r][0] - b.s * S.G[r][1];
      fb[r] = b.s * S.G[r][0] + b.c * S.G[r][1];
    }
    double S5[15];
    {
      double s00 = DELTA_C_LOCAL + isd + iel, s01 = 0, s02 = 0, s03 = 0, s04 = 0;
      double s11 = DELTA_C_LOCAL, s12 = 0, s22 = DELTA_C_LOCAL, s33 = DELTA_C_LOCAL, s34 = 0, s44 = DELTA_C_LOCAL;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        s00 += B.ba[r] * B.ba[r] * sl[r] + B.bb[r] * B.bb[r] * smu[r];
        s01 -= B.ba[r] * ea[r] * sl[r], s02 -= B.ba[r] * fa[r] * sl[r];
        s03 -= B.bb[r] * eb[r] * smu[r], s04 -= B.bb[r] * fb[r] * smu[r];
        s11 += ea[r] * ea[r] * sl[r], s12 += ea[r] * fa[r] * sl[r], s22 += fa[r] * fa[r] * sl[r];
        s33 += eb[r] * eb[r] * smu[r], s34 += eb[r] * fb[r] * smu[r], s44 += fb[r] * fb[r] * smu[r];
      }
      S5[sym(0, 0)] = s00, S5[sym(1, 0)] = s01, S5[sym(2, 0)] = s02, S5[sym(3, 0)] = s03, S5[sym(4, 0)] = s04;
      S5[sym(1, 1)] = s11, S5[sym(2, 1)] = s12, S5[sym(2, 2)] = s22, S5[sym(3, 1)] = 0, S5[sym(3, 2)] = 0, S5[sym(4, 1)] = 0, S5[sym(4, 2)] = 0;
      S5[sym(3, 3)] = s33, S5[sym(4, 3)] = s34, S5[sym(4, 4)] = s44;
    }
    if (!chol_packed<5>(S5)) *ok = 0;
    // Z = S6^-1 E' with E = [[0, 1, 0, -1, 0, -2 s0], [0, 0, 1, 0, -1, -2 s1]]
    double Z0[6] = {0, 1, 0, -1, 0, -2.0 * B.s[0] * idn}, Z1[6] = {0, 0, 1, 0, -1, -2.0 * B.s[1] * idn};
    chol_solve_packed<5>(S5, Z0);
    chol_solve_packed<5>(S5, Z1);
    // Ms = diag(hs) + E Z  (2 x 2, symmetric positive definite)
    double m00 = hs0 + Z0[1] - Z0[3] - 2.0 * B.s[0] * Z0[5];
    double m01 = Z1[1] - Z1[3] - 2.0 * B.s[0] * Z1[5];
    double m11 = hs1 + Z1[2] - Z1[4] - 2.0 * B.s[1] * Z1[5];
    double det = m00 * m11 - m01 * m01;
    if (!(m00 > 0) || !(det > 0)) *ok = 0;
    const double idet = 1.0 / det;
    // coupling columns: (x_a, y_a, psi_a, x_b, y_b, psi_b)
    double dRua[2] = {-a.s * B.ua[0] - a.c * B.ua[1], a.c * B.ua[0] - a.s * B.ua[1]};
    double dRub[2] = {-b.s * B.ub[0] - b.c * B.ub[1], b.c * B.ub[0] - b.s * B.ub[1]};
    double dRtea[2] = {-a.s * ye1[0] + a.c * ye1[1], -a.c * ye1[0] - a.s * ye1[1]};  // (dR/dpsi)' ye1
    double dRteb[2] = {-b.s * ye2[0] + b.c * ye2[1], -b.c * ye2[0] - b.s * ye2[1]};
    // Coupling columns C (16 x 6, sparse): column k < 3 (pose a) touches lam (rows 0-3), yd (row 8) and, for psi, ye1 (9, 10);
    // column k >= 3 (pose b) touches mu (rows 4-7), yd and ye2 (11, 12).  The entries are kept in registers (no local arrays:
    // the per-thread stack lives in L2 at this shared-memory carve-out) and the Schur complement C'X is accumulated column by
    // column while the solves X[:, k] are produced and streamed to global memory.
    double ca[3][4], cb[3][4];  // lam rows of columns 0-2, mu rows of columns 3-5
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const double dax = -S.G[r][0] * a.s - S.G[r][1] * a.c, day = S.G[r][0] * a.c - S.G[r][1] * a.s;
      const double dbx = -S.G[r][0] * b.s - S.G[r][1] * b.c, dby = S.G[r][0] * b.c - S.G[r][1] * b.s;
      ca[0][r] = -yd * ea[r], ca[1][r] = -yd * fa[r];
      ca[2][r] = -yd * (dax * a.x + day * a.y) + S.G[r][0] * dRtea[0] + S.G[r][1] * dRtea[1];
      cb[0][r] = -yd * eb[r], cb[1][r] = -yd * fb[r];
      cb[2][r] = -yd * (dbx * b.x + dby * b.y) + S.G[r][0] * dRteb[0] + S.G[r][1] * dRteb[1];
    }
    const double c8[6] = {-B.Rua[0], -B.Rua[1], -(a.x * dRua[0] + a.y * dRua[1]), -B.Rub[0], -B.Rub[1], -(b.x * dRub[0] + b.y * dRub[1])};
    // block solves, node minor ([P][112][Mv]): the threads of a warp own consecutive nodes, every store / load is one coalesced line
    double* xp = W.XP + (size_t)p * 112 * L.Mv + n;
    const int xs = L.Mv;
    double CX[6][7];  // C' X
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double bl[4] = {0, 0, 0, 0}, bm[4] = {0, 0, 0, 0}, by[6] = {0, 0, 0, 0, 0, 0}, bs[2] = {0, 0};
      if (k < 3) {
#pragma unroll
        for (int r = 0; r < 4; ++r) bl[r] = ca[k][r];
        by[0] = c8[k];
        if (k == 2) by[1] = dRua[0], by[2] = dRua[1];
      } else if (k < 6) {
#pragma unroll
        for (int r = 0; r < 4; ++r) bm[r] = cb[k - 3][r];
        by[0] = c8[k];
        if (k == 5) by[3] = dRub[0], by[4] = dRub[1];
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) bl[r] = -W.gphi[L.PL(p, r, n)], bm[r] = -W.gphi[L.PM(p, r, n)];
        // the row residuals are read from W.c (not recomputed): the refinement solve passes its own right-hand side
        by[0] = -W.c[L.YPAIR(p, 0, n)] - W.gphi[L.PSD(p, n)] * isd + W.gphi[L.PEL(p, n)] * iel;
#pragma unroll
        for (int r = 1; r < 5; ++r) by[r] = -W.c[L.YPAIR(p, r, n)];
        by[5] = -W.c[L.YPAIR(p, 5, n)] - W.gphi[L.PSN(p, n)] * isn;
        bs[0] = -W.gphi[L.PS(p, 0, n)], bs[1] = -W.gphi[L.PS(p, 1, n)];
      }
      // ry = J D^-1 b - by   (J rows: yd [-ba, -bb]; e1 [ea; fa | 0]; e2 [0 | eb; fb]; yn 0)
      double ry[6] = {-by[0], -by[1], -by[2], -by[3], -by[4], -by[5]};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double tl = bl[r] * sl[r], tm = bm[r] * smu[r];
        ry[0] -= B.ba[r] * tl + B.bb[r] * tm;
        ry[1] += ea[r] * tl, ry[2] += fa[r] * tl, ry[3] += eb[r] * tm, ry[4] += fb[r] * tm;
      }
      // u = S6^-1 ry ; ds = Ms^-1 (bs - E u) ; dy = u + Z ds
      ry[5] *= idn;
      chol_solve_packed<5>(S5, ry);
      double q0 = bs[0] - (ry[1] - ry[3] - 2.0 * B.s[0] * ry[5]);
      double q1 = bs[1] - (ry[2] - ry[4] - 2.0 * B.s[1] * ry[5]);
      double ds0 = (m11 * q0 - m01 * q1) * idet, ds1 = (m00 * q1 - m01 * q0) * idet;
      double dy[6], xl[4], xm[4];
#pragma unroll
      for (int r = 0; r < 6; ++r) dy[r] = ry[r] + Z0[r] * ds0 + Z1[r] * ds1;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        xl[r] = (bl[r] - (-B.ba[r] * dy[0] + ea[r] * dy[1] + fa[r] * dy[2])) * sl[r];
        xm[r] = (bm[r] - (-B.bb[r] * dy[0] + eb[r] * dy[3] + fb[r] * dy[4])) * smu[r];
        xp[(r * 7 + k) * xs] = xl[r], xp[((4 + r) * 7 + k) * xs] = xm[r];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) xp[((8 + r) * 7 + k) * xs] = dy[r];
      xp[(14 * 7 + k) * xs] = ds0, xp[(15 * 7 + k) * xs] = ds1;
      // row r of C' X[:, k]
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        double sa_ = c8[r] * dy[0], sb_ = c8[3 + r] * dy[0];
#pragma unroll
        for (int m = 0; m < 4; ++m) sa_ += ca[r][m] * xl[m], sb_ += cb[r][m] * xm[m];
        CX[r][k] = sa_, CX[3 + r][k] = sb_;
      }
      CX[2][k] += dRua[0] * dy[1] + dRua[1] * dy[2];
      CX[5][k] += dRub[0] * dy[3] + dRub[1] * dy[4];
    }
    // Schur complement on (pose_a, pose_b): direct Hessian - C' Xc ; gradient C' Xr
    double* ph = W.PH + (size_t)p * 27 * L.Mv + n;  // [P][27][Mv], node minor
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        double h = 0;
        if (r == 2 && q == 0) h = -yd * dRua[0];
        if (r == 2 && q == 1) h = -yd * dRua[1];
        if (r == 2 && q == 2) h = yd * (a.x * B.Rua[0] + a.y * B.Rua[1]) - (ye1[0] * B.Rua[0] + ye1[1] * B.Rua[1]);
        if (r == 5 && q == 3) h = -yd * dRub[0];
        if (r == 5 && q == 4) h = -yd * dRub[1];
        if (r == 5 && q == 5) h = yd * (b.x * B.Rub[0] + b.y * B.Rub[1]) - (ye2[0] * B.Rub[0] + ye2[1] * B.Rub[1]);
        ph[sym(r, q) * xs] = h - CX[r][q];
      }
      ph[(21 + r) * xs] = CX[r][6];
    }
  }
}

OBCA_HDN void pair_eliminate(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, int* ok) {
  assume_scratch(W);
  OBCA_ASSUME_STATIC(L, S);
  for (int it = ctx.tid; it < L.nPairNodes; it += ctx.nt) {  // compact index over the existing (pair, node) blocks
    int p = 0, n = it;
    while (n >= L.Mp[p]) n -= L.Mp[p], ++p;
    Pose a, b;
    load_pose(L, W.x, L.pa[p], n, a);
    load_pose(L, W.x, L.pb[p], n, b);
    pair_block_eliminate(L, S, W, p, n, a, b, ok);
  }
}

// one (node, obstacle) block: unknown order lam 0-3, mu 4-7 | y1 8, y2 9-10, y3 11 (sd, el folded into the y1 pivot);
// stores the block solves for the back-substitution and adds the Schur complement to the node Hessian H / gradient g
OBCA_HD void obs_block_eliminate(const Lay& L, const Stat& S, const Scratch& W, int a, int n, int j, const Pose& p, double* H, double* g, int* ok) {
  const double *x = W.x, *y = W.y;
  {
      ObsBlk B;
      for (int r = 0; r < 4; ++r) B.lam[r] = x[L.LAM(a, j, r, n)], B.mu[r] = x[L.MU(a, j, r, n)];
      B.sd = x[L.SD(a, j, n)];
      B.el = x[L.EL(a, j, n)];
      obs_residual(S, j, p, B);
      double y1 = y[L.YOBS(a, j, 0, n)], y2[2] = {y[L.YOBS(a, j, 1, n)], y[L.YOBS(a, j, 2, n)]}, y3 = y[L.YOBS(a, j, 3, n)];
      double y3p = y3 > 0 ? y3 : 0.0;  // local convexification: exact at KKT points (y3 >= 0)
      // Block elimination in registers: lam (H = diag + 2 y3+ A A'), mu (diagonal), then the 4 x 4 Schur complement
      //   S = D + Jl Hl^-1 Jl' + Jm Dm^-1 Jm'  on (y1, y2, y3); 4 right-hand sides (3 pose couplings + residual).
      double* xo = W.XO + (size_t)(a * L.O + j) * 48 * L.Mv + n;  // node minor ([V][O][48][Mv]): coalesced across the threads of a warp
      const int xs = L.Mv;
      double CX[3][4];  // C' X
      double dRy[2] = {-p.s * y2[0] - p.c * y2[1], p.c * y2[0] - p.s * y2[1]};   // (dR/dpsi) y2
      double dRtu[2] = {-p.s * B.u[0] + p.c * B.u[1], -p.c * B.u[0] - p.s * B.u[1]};  // (dR'/dpsi) u
      double Hl[10], Jl[4][4], sm[4], Cl[4][3], bl[4], bm[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double* A = S.obsA[j][r];
#pragma unroll
        for (int q = 0; q <= r; ++q) Hl[sym(r, q)] = 2.0 * y3p * (A[0] * S.obsA[j][q][0] + A[1] * S.obsA[j][q][1]);
        Hl[sym(r, r)] += W.sig[L.LAM(a, j, r, n)];
        sm[r] = rcp_pos(W.sig[L.MU(a, j, r, n)]);  // inverse
        Jl[0][r] = B.Atb[r];
        Jl[1][r] = p.c * A[0] + p.s * A[1];
        Jl[2][r] = -p.s * A[0] + p.c * A[1];
        Jl[3][r] = 2.0 * (A[0] * B.u[0] + A[1] * B.u[1]);
        Cl[r][0] = y1 * A[0];
        Cl[r][1] = y1 * A[1];
        Cl[r][2] = A[0] * dRy[0] + A[1] * dRy[1];
        bl[r] = -W.gphi[L.LAM(a, j, r, n)];
        bm[r] = -W.gphi[L.MU(a, j, r, n)];
      }
      const double isd = rcp_pos(W.sig[L.SD(a, j, n)]), iel = rcp_pos(W.sig[L.EL(a, j, n)]);
      if (!chol_packed<4>(Hl)) *ok = 0;
      double Wl[4][4];  // Wl[i] = Hl^-1 Jl[i]'
#pragma unroll
      for (int i2 = 0; i2 < 4; ++i2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) Wl[i2][r] = Jl[i2][r];
        chol_solve_packed<4>(Hl, Wl[i2]);
      }
      // Jm rows: -g, G[:,0], G[:,1], 0
      double Ss[10];
#pragma unroll
      for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
        for (int j2 = 0; j2 <= i2; ++j2) {
          double acc = 0;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc += Jl[i2][r] * Wl[j2][r];
            double ji = i2 == 0 ? -S.g[r] : (i2 == 1 ? S.G[r][0] : (i2 == 2 ? S.G[r][1] : 0.0));
            double jj = j2 == 0 ? -S.g[r] : (j2 == 1 ? S.G[r][0] : (j2 == 2 ? S.G[r][1] : 0.0));
            acc += ji * jj * sm[r];
          }
          Ss[sym(i2, j2)] = acc;
        }
      Ss[sym(0, 0)] += DELTA_C_LOCAL + isd + iel;
      Ss[sym(1, 1)] += DELTA_C_LOCAL, Ss[sym(2, 2)] += DELTA_C_LOCAL, Ss[sym(3, 3)] += DELTA_C_LOCAL;
      if (!chol_packed<4>(Ss)) *ok = 0;
      const double Cy[4][3] = {{B.u[0], B.u[1], 0.0}, {0.0, 0.0, dRtu[0]}, {0.0, 0.0, dRtu[1]}, {0.0, 0.0, 0.0}};
      const double byr[4] = {-W.c[L.YOBS(a, j, 0, n)] - W.gphi[L.SD(a, j, n)] * isd + W.gphi[L.EL(a, j, n)] * iel, -W.c[L.YOBS(a, j, 1, n)],
                             -W.c[L.YOBS(a, j, 2, n)], -W.c[L.YOBS(a, j, 3, n)]};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double t[4], ry[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) t[r] = k < 3 ? Cl[r][k] : bl[r];
        chol_solve_packed<4>(Hl, t);
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          double acc = k < 3 ? -Cy[i2][k] : -byr[i2];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc += Jl[i2][r] * t[r];
            if (k == 3) {
              double ji = i2 == 0 ? -S.g[r] : (i2 == 1 ? S.G[r][0] : (i2 == 2 ? S.G[r][1] : 0.0));
              acc += ji * bm[r] * sm[r];
            }
          }
          ry[i2] = acc;
        }
        chol_solve_packed<4>(Ss, ry);
        double xl[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          double dl = t[r], jm = -S.g[r] * ry[0] + S.G[r][0] * ry[1] + S.G[r][1] * ry[2];
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) dl -= Wl[i2][r] * ry[i2];
          xl[r] = dl;
          // block solves go straight to global memory (back-substitution); no per-thread local array
          xo[(r * 4 + k) * xs] = dl;
          xo[((4 + r) * 4 + k) * xs] = ((k == 3 ? bm[r] : 0.0) - jm) * sm[r];
          xo[((8 + r) * 4 + k) * xs] = ry[r];
        }
        // column k of C'X: the lam rows carry Cl, the y rows carry Cy (the mu rows do not couple to the pose)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          double sacc = Cy[0][r] * ry[0] + Cy[1][r] * ry[1] + Cy[2][r] * ry[2];
#pragma unroll
          for (int m = 0; m < 4; ++m) sacc += Cl[m][r] * xl[m];
          CX[r][k] = sacc;
        }
      }
      H[sym(2, 2)] -= y2[0] * (p.c * B.u[0] + p.s * B.u[1]) + y2[1] * (-p.s * B.u[0] + p.c * B.u[1]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int q = 0; q <= r; ++q) H[sym(r, q)] -= CX[r][q];
        g[r] += CX[r][3];
      }
  }
}

OBCA_HD void obs_block_backsub(const Lay& L, const Scratch& W, int a, int n, int j, const double* dp) {
  const double* xo = W.XO + (size_t)(a * L.O + j) * 48 * L.Mv + n;
  const int xs = L.Mv;
  double r[12];
  for (int m = 0; m < 12; ++m) r[m] = xo[(m * 4 + 3) * xs] - xo[(m * 4 + 0) * xs] * dp[0] - xo[(m * 4 + 1) * xs] * dp[1] - xo[(m * 4 + 2) * xs] * dp[2];
  for (int q = 0; q < 4; ++q) {
    W.dx[L.LAM(a, j, q, n)] = r[q];
    W.dx[L.MU(a, j, q, n)] = r[4 + q];
    W.dy[L.YOBS(a, j, q, n)] = r[8 + q];
  }
  // sd: sig dsd - dy1 = -gphi ; el: sig del + dy1 = -gphi
  W.dx[L.SD(a, j, n)] = (r[8] - W.gphi[L.SD(a, j, n)]) * rcp_pos(W.sig[L.SD(a, j, n)]);
  W.dx[L.EL(a, j, n)] = -(r[8] + W.gphi[L.EL(a, j, n)]) * rcp_pos(W.sig[L.EL(a, j, n)]);
}

OBCA_HD void pair_block_backsub(const Lay& L, const Scratch& W, int p, int n, const double* dp) {
  const double* xp = W.XP + (size_t)p * 112 * L.Mv + n;
  const int xs = L.Mv;
  double r[16];
  for (int m = 0; m < 16; ++m) {
    double s = xp[(m * 7 + 6) * xs];
    for (int q = 0; q < 6; ++q) s -= xp[(m * 7 + q) * xs] * dp[q];
    r[m] = s;
  }
  for (int q = 0; q < 4; ++q) W.dx[L.PL(p, q, n)] = r[q], W.dx[L.PM(p, q, n)] = r[4 + q];
  for (int q = 0; q < 6; ++q) W.dy[L.YPAIR(p, q, n)] = r[8 + q];
  W.dx[L.PS(p, 0, n)] = r[14];
  W.dx[L.PS(p, 1, n)] = r[15];
  W.dx[L.PSD(p, n)] = (r[8] - W.gphi[L.PSD(p, n)]) * rcp_pos(W.sig[L.PSD(p, n)]);
  W.dx[L.PEL(p, n)] = -(r[8] + W.gphi[L.PEL(p, n)]) * rcp_pos(W.sig[L.PEL(p, n)]);
  W.dx[L.PSN(p, n)] = (r[13] - W.gphi[L.PSN(p, n)]) * rcp_pos(W.sig[L.PSN(p, n)]);
}

// ------------------------------------------------------------------------------------------------
// [LOCAL] node assembly: obstacle / tube elimination, bounds, cost and collocation curvature
// ------------------------------------------------------------------------------------------------
OBCA_HDN void node_assemble(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, int* ok, double* hdtdt_out) {
  assume_scratch(W);
  OBCA_ASSUME_STATIC(L, S);
  const double *x = W.x, *y = W.y;
  const double dt = x[L.oDT], idt = 1.0 / dt;
  double hdt_part = 0;
  for (int it = ctx.tid; it < L.V * L.Mv; it += ctx.nt) {
    int a = it / L.Mv, n = it % L.Mv;
    if (n >= L.M[a]) continue;
    int i = n / NK, k = n % NK, n0 = i * NK;
    double z[NZ];
    for (int q = 0; q < NZ; ++q) z[q] = x[L.Z(a, q, n)];
    Pose p = {z[0], z[1], z[2], cos(z[2]), sin(z[2])};
    double v = z[3], de = z[4], ua = z[5], uw = z[6];
    double tde = tan(de), sec2 = 1.0 + tde * tde;
    double H[28], g[NZ], hd[NZ];
    for (int q = 0; q < 28; ++q) H[q] = 0;
    for (int q = 0; q < NZ; ++q) {
      H[sym(q, q)] = W.sig[L.Z(a, q, n)];
      g[q] = W.gphi[L.Z(a, q, n)];
      hd[q] = 0;
    }
    double bk = S.cB[k], bdt = bk * dt;
    H[sym(3, 3)] += bdt * 2.0 * uw * uw;
    H[sym(6, 6)] += bdt * 2.0 * v * v;
    H[sym(6, 3)] += bdt * 4.0 * v * uw;
    H[sym(4, 4)] += bdt * 2.0;
    H[sym(5, 5)] += bdt * 2.0;
    hd[3] = bk * 2.0 * v * uw * uw;
    hd[4] = bk * 2.0 * de;
    hd[5] = bk * 2.0 * ua;
    hd[6] = bk * 2.0 * v * v * uw;
    double yc[5];
    for (int q = 0; q < 5; ++q) {
      yc[q] = y[L.YCOL(a, q, n)];
      double s = 0, pl = 0;
      for (int kk = 0; kk < NK; ++kk) s += S.cA[k][kk] * y[L.YCOL(a, q, n0 + kk)];
      for (int j = 0; j < NK; ++j) pl += S.cA[j][k] * x[L.Z(a, q, n0 + j)];
      hd[q] -= s * idt * idt;
      hdt_part += yc[q] * 2.0 * pl * idt * idt * idt;
    }
    H[sym(2, 2)] += yc[0] * v * p.c + yc[1] * v * p.s;
    H[sym(3, 2)] += yc[0] * p.s - yc[1] * p.c;
    H[sym(4, 3)] -= yc[2] * sec2 / S.wb;
    H[sym(4, 4)] -= yc[2] * 2.0 * v * sec2 * tde / S.wb;
    for (int j = 0; j < L.O; ++j) obs_block_eliminate(L, S, W, a, n, j, p, H, g, ok);
    // ---- tube set (slack and multiplier eliminated analytically)
    int q = tube_set_at(L, a, n);
    if (q >= 1) {
      for (int r = 0; r < 8; ++r) {
        const double* t = S.tube_row(L, a, q, r / 4, r % 4);
        double gr[3] = {t[0], t[1], r < 4 ? 0.0 : S.wb * (-t[0] * p.s + t[1] * p.c)};
        double sg = W.sig[L.TS(a, q - 1, r)];
        double w = sg * W.c[L.YTUBE(a, q - 1, r)] + W.gphi[L.TS(a, q - 1, r)];
        for (int m = 0; m < 3; ++m) {
          g[m] -= gr[m] * w;
          for (int mm = 0; mm <= m; ++mm) H[sym(m, mm)] += sg * gr[m] * gr[mm];
        }
        if (r >= 4) H[sym(2, 2)] += y[L.YTUBE(a, q - 1, r)] * S.wb * (t[0] * p.c + t[1] * p.s);
      }
    }
    // ---- pair Schur complements (diagonal blocks and gradients)
    for (int pp = 0; pp < L.P; ++pp) {
      if (n >= L.Mp[pp]) continue;
      const double* ph = W.PH + (size_t)pp * 27 * L.Mv + n;
      int off = L.pa[pp] == a ? 0 : (L.pb[pp] == a ? 3 : -1);
      if (off < 0) continue;
      for (int r = 0; r < 3; ++r) {
        for (int m = 0; m <= r; ++m) H[sym(r, m)] += ph[sym(off + r, off + m) * L.Mv];
        g[r] += ph[(21 + off + r) * L.Mv];
      }
    }
    // node minor ([V][28][Mv], [V][7][Mv]): the threads of a warp own consecutive nodes, every store is one coalesced line
    double* hn = W.HN + (size_t)a * 28 * L.Mv + n;
    double* gn = W.GN + (size_t)a * 7 * L.Mv + n;
    double* hdn = W.HD + (size_t)a * 7 * L.Mv + n;
    for (int m = 0; m < 28; ++m) hn[(size_t)m * L.Mv] = H[m];
    for (int m = 0; m < NZ; ++m) gn[(size_t)m * L.Mv] = g[m], hdn[(size_t)m * L.Mv] = hd[m];
  }
  double hdt = cta_sum(ctx, hdt_part);
  for (int a = 0; a < L.V; ++a) hdt += 2.0 * L.N[a] * L.N[a];
  *hdtdt_out = hdt + W.sig[L.oDT];
}

// ------------------------------------------------------------------------------------------------
// [NULLSP] one warp per (vehicle, interval)
//
// G_w (nr x 35) = Jacobian of the interval's collocation (+ terminal + implied) rows w.r.t. the stage variables of
// nodes 1..K.  The controls (a_k, w_k) and the poses (x, y, psi) of the interior nodes are eliminated exactly through
// their defining rows (nullspace_block); what remains is a Householder QR of the node-0 / terminal / implied rows on
// the 10 variables (v_j, delta_j), with a rank test: where LICQ fails (a vehicle standing still over a whole interval
// makes the over-collocated rows dependent) the dependent rows are dropped and the null space grows, the analogue of
// IPOPT's delta_c perturbation for a singular Jacobian.
// QR record: [NQ][NC] rows 0..9 = reflectors below / R on and above the staircase, rows 10..24 = W (the multipliers of
// the eliminated (x, y, psi) rows onto every remaining row), [35] tau, [35] pivot column of every staircase row,
// [4] (rank, nr, ndrop, -), dropped-row records, nu, [2][NC] control coefficients of the terminal / implied rows.
// apply Q = H_0 ... H_{rk-1} (transpose = false) or Q' (transpose = true) to v[NU2]
OBCA_HD void apply_q(const double* QRm, int rk, double* v, bool transpose) {
  const double* tau = QRm + QR_TAU;
  const double* piv = QRm + QR_PIV;
  for (int jj = 0; jj < rk; ++jj) {
    int i = transpose ? jj : rk - 1 - jj;          // reflector i acts on rows i..34, stored in column piv[i]
    int col = (int)piv[i];
    double s = v[i];
    for (int r = i + 1; r < NU2; ++r) s += QRm[r * NC + col] * v[r];
    s *= tau[i];
    v[i] -= s;
    for (int r = i + 1; r < NU2; ++r) v[r] -= s * QRm[r * NC + col];
  }
}

// ---- warp-cooperative execution helpers: on the device the body runs once per lane; the host emulation loops over lanes
#if defined(__CUDA_ARCH__)
#define OBCA_LANES(lane) for (int lane = (ctx.tid & 31), once_ = 1; once_; once_ = 0)
#define OBCA_WARP_SYNC() __syncwarp()
#else
#define OBCA_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define OBCA_WARP_SYNC()
#endif

#if defined(OBCA_HOST_EMU)
#define OBCA_DBG(...) do { if (getenv("OBCA_TRACE")) printf(__VA_ARGS__); } while (0)
#else
#define OBCA_DBG(...) do { } while (0)
#endif
constexpr int NSW = 2720;  // shared-memory doubles per warp of the null-space phase

// apply Q or Q' to the strided vector v[q * stride], q < 35 (reflectors in the shared-memory QR matrix)
OBCA_HD void apply_q_strided(const double* __restrict__ Mq, const double* __restrict__ tau, const double* __restrict__ piv, int rk,
                             double* __restrict__ v, int stride, bool transpose) {
  OBCA_ASSUME_SHARED(Mq);
  OBCA_ASSUME_SHARED(tau);
  OBCA_ASSUME_SHARED(piv);
  OBCA_ASSUME_SHARED(v);
  for (int jj = 0; jj < rk; ++jj) {
    const int i = transpose ? jj : rk - 1 - jj;
    const double* __restrict__ u = Mq + (int)piv[i];
    double s0 = v[i * stride], s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = i + 1;
    for (; r + 3 < NW; r += 4) {  // four accumulators: loads issue back to back, the dependent FP64 chain is a quarter as long
      const double u0 = u[r * NC], u1 = u[(r + 1) * NC], u2 = u[(r + 2) * NC], u3 = u[(r + 3) * NC];
      const double v0 = v[r * stride], v1 = v[(r + 1) * stride], v2 = v[(r + 2) * stride], v3 = v[(r + 3) * stride];
      s0 += u0 * v0, s1 += u1 * v1, s2 += u2 * v2, s3 += u3 * v3;
    }
    for (; r < NW; ++r) s0 += u[r * NC] * v[r * stride];
    const double s = ((s0 + s1) + (s2 + s3)) * tau[i];
    v[i * stride] -= s;
    r = i + 1;
    for (; r + 3 < NW; r += 4) {  // loads first, then the stores: no store -> load serialisation
      const double u0 = u[r * NC], u1 = u[(r + 1) * NC], u2 = u[(r + 2) * NC], u3 = u[(r + 3) * NC];
      const double v0 = v[r * stride], v1 = v[(r + 1) * stride], v2 = v[(r + 2) * stride], v3 = v[(r + 3) * stride];
      v[r * stride] = v0 - s * u0, v[(r + 1) * stride] = v1 - s * u1, v[(r + 2) * stride] = v2 - s * u2, v[(r + 3) * stride] = v3 - s * u3;
    }
    for (; r < NW; ++r) v[r * stride] -= s * u[r * NC];
  }
}

// One (vehicle, interval) block, processed by one warp with its working set in shared memory `sw` (NSW doubles):
// rows = 30 collocation rows (+ terminal rows in the last interval) + the implied rows `ex` received from interval
// i+1 (they act on the node-K variables).  Writes T, s0, the QR record, the projected Hessian, and the implied rows
// `em` for interval i-1.
OBCA_HDN void nullspace_block(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, int a, int i, const double* ex, double* em,
                              int* ok, double* sw) {
  OBCA_ASSUME_SHARED(sw);
  OBCA_ASSUME_STATIC(L, S);
  assume_scratch(W);
  OBCA_ASSUME_GLOBAL(ex), OBCA_ASSUME_GLOBAL(em);
  const double* x = W.x;
  const double dt = x[L.oDT], idt = 1.0 / dt;
  const int n0 = i * NK;
  const bool last = (i == L.N[a] - 1);
  const int nterm = last ? (4 + L.heading[a]) : 0;
  const int nex = (int)ex[0];
  const int nr = 30 + nterm + nex;
  double* Mq = sw;                 // [NQ][NC] (later reused for the projection: 42 x NRED + ...)
  double* G0 = Mq + NQ * NC;       // [NC][7]
  double* gd = G0 + NC * 7;        // [NC]
  double* rr = gd + NC;            // [NC]
  double* tau = rr + NC;           // [35]
  double* piv = tau + NW;          // [35]
  double* T = piv + NW;            // [35][NRED]
  double* s0 = T + NW * NRED;      // [35]
  double* zb = s0 + NW;            // [42] states of the interval
  double* hdv = zb + NS;           // [42] node x dt cross Hessian
  double* gnv = hdv + NS;          // [42] node gradients
  double* wred = gnv + NS;         // [64] reduction scratch
  double* misc = wred + 64;        // [64]: al[35], h[9], flags
  double* bu = misc + 64;          // [2][NC] coefficients of (a_K, w_K) in the rows before the elimination
  double* ndv = bu + 2 * NC;       // [5][5] per interior node: cos, sin, tan/wb, v sec^2/wb, v
  double* rref = ndv + 25;         // [NC] norm of every remaining row before the elimination
  double* QRg = W.QR + (size_t)(a * L.Nmax + i) * QRSZ;
  OBCA_LANES(lane) {
    for (int q = lane; q < NQ * NC + NC * 7; q += 32) Mq[q] = 0;  // Mq and G0 are contiguous
    for (int q = lane; q < 2 * NC; q += 32) bu[q] = 0;
    for (int q = lane; q < NS; q += 32) {
      zb[q] = x[L.Z(a, q % NZ, n0 + q / NZ)];
      hdv[q] = W.HD[((size_t)a * 7 + q % NZ) * L.Mv + n0 + q / NZ];
      gnv[q] = W.GN[((size_t)a * 7 + q % NZ) * L.Mv + n0 + q / NZ];
    }
  }
  OBCA_WARP_SYNC();
  // variable rows of the matrix: (v_j, delta_j) -> (j-1)*2 + (q-3) in 0..9 (the QR part), (x, y, psi)_j -> 10 + q*5 + (j-1)
#define OBCA_VROW(j, q) ((q) >= 3 ? ((j) - 1) * 2 + (q) - 3 : 10 + (q) * 5 + (j) - 1)
  OBCA_LANES(lane) {
    if (lane < 30) {
      int k = lane / 5, q = lane % 5, r = lane;
      double psi = zb[k * NZ + 2], v = zb[k * NZ + 3], de = zb[k * NZ + 4];
      double cs = cos(psi), sn = sin(psi), tde = tan(de), sec2 = 1.0 + tde * tde;
      double pl = 0;
      for (int j = 0; j < NK; ++j) pl += S.cA[j][k] * zb[j * NZ + q];
      G0[r * 7 + q] += S.cA[0][k] * idt;
      gd[r] = -pl * idt * idt;
      rr[r] = W.c[L.YCOL(a, q, n0 + k)];
      if (k == 0) {
        // rows of node 0: collocation coefficients on the stage variables, dynamics derivatives on xi
        for (int j = 1; j < NK; ++j) Mq[OBCA_VROW(j, q) * NC + r] = S.cA[j][0] * idt;
        if (q == 0) G0[r * 7 + 2] -= -v * sn, G0[r * 7 + 3] -= cs;
        else if (q == 1) G0[r * 7 + 2] -= v * cs, G0[r * 7 + 3] -= sn;
        else if (q == 2) G0[r * 7 + 3] -= tde / S.wb, G0[r * 7 + 4] -= v * sec2 / S.wb;
        else if (q == 3) G0[r * 7 + 5] -= 1.0;
        else G0[r * 7 + 6] -= 1.0;
      } else if (q == 0) {
        // rows (k >= 1, q) are the defining rows of (x, y, psi, a, w)_k and are never stored; node data for their elimination
        double* nd = ndv + (k - 1) * 5;
        nd[0] = cs, nd[1] = sn, nd[2] = tde / S.wb, nd[3] = v * sec2 / S.wb, nd[4] = v;
      }
    } else if (lane == 30) {
      int r = 30;
      if (last) {
        if (L.heading[a]) {
          Mq[OBCA_VROW(5, 2) * NC + r] = 1.0, gd[r] = 0, rr[r] = W.c[L.YTERM(a, 0)];
          ++r;
        }
        for (int m = 3; m < NZ; ++m, ++r) {
          if (m < 5) Mq[OBCA_VROW(5, m) * NC + r] = 1.0;
          else bu[(m - 5) * NC + r] = 1.0;
          gd[r] = 0, rr[r] = W.c[L.YTERM(a, m - 2)];
        }
      }
      for (int e = 0; e < nex; ++e, ++r) {
        const double* h = ex + 1 + e * 9;
        for (int m = 0; m < 5; ++m) Mq[OBCA_VROW(5, m) * NC + r] = h[m];
        bu[r] = h[5], bu[NC + r] = h[6];
        gd[r] = h[7], rr[r] = h[8];
      }
    }
  }
  OBCA_WARP_SYNC();
  // Exact eliminations, one lane per remaining row r (rows of node 0, terminal rows, implied rows):
  //  (1) controls: row (k, 3) reads -a_k + sum_j cA[j][k]/dt v_j + ... = -r (row (k, 4): w_k, delta_j), so a_k, w_k are affine in
  //      the states; they appear in the terminal / implied rows of node K, where they are substituted;
  //  (2) (x, y, psi): rows (k, 0..2), k >= 1, have the constant interior collocation block C = cA'/dt on them (plus the
  //      psi_k terms of the x and y rows), B = [[C, 0, Dx], [0, C, Dy], [0, 0, C]]; with W = M B^-1 the rows become
  //      K - W N on (v, delta).  W overwrites M (it is needed again for the multipliers).
  // The QR below then sees 10 variables and 5 (+ terminal + implied) rows, and a rank deficiency (standing vehicle) can only
  // show up there: the eliminated blocks are nonsingular whatever the state.
  OBCA_LANES(lane) {
    const int r = lane < 5 ? lane : 30 + lane - 5;
    if (r < nr && lane < 5 + NC - 30) {
      if (r >= 30) {
        const double ba = bu[r], bw = bu[NC + r];
        for (int j = 1; j < NK; ++j) {
          const double coef = S.cA[j][5] * idt;
          Mq[OBCA_VROW(j, 3) * NC + r] += ba * coef;
          Mq[OBCA_VROW(j, 4) * NC + r] += bw * coef;
        }
        G0[r * 7 + 3] += ba * S.cA[0][5] * idt, G0[r * 7 + 4] += bw * S.cA[0][5] * idt;
        gd[r] += ba * gd[28] + bw * gd[29];
        rr[r] += ba * rr[28] + bw * rr[29];
      }
      // scale of the row before the elimination: the rank test below must not accept a row that cancelled to rounding noise
      {
        double sq = 0;
        for (int q = 0; q < NQ; ++q) sq += Mq[q * NC + r] * Mq[q * NC + r];
        rref[r] = sqrt(sq);
      }
      double wx[5], wy[5], wp[5], t[5];
      for (int k = 0; k < 5; ++k) {
        double sx = 0, sy = 0;
        for (int j = 0; j < 5; ++j) sx += Mq[(10 + j) * NC + r] * S.cAi[j][k], sy += Mq[(15 + j) * NC + r] * S.cAi[j][k];
        wx[k] = sx * dt, wy[k] = sy * dt;
      }
      for (int j = 0; j < 5; ++j) {
        const double* nd = ndv + j * 5;  // Dx = v sin(psi), Dy = -v cos(psi)
        t[j] = Mq[(20 + j) * NC + r] - wx[j] * nd[4] * nd[1] + wy[j] * nd[4] * nd[0];
      }
      for (int k = 0; k < 5; ++k) {
        double sp = 0;
        for (int j = 0; j < 5; ++j) sp += t[j] * S.cAi[j][k];
        wp[k] = sp * dt;
      }
      double g0 = 0, g1 = 0, g2 = 0, sgd = 0, srr = 0;
      for (int k = 0; k < 5; ++k) {
        const double* nd = ndv + k * 5;
        Mq[(10 + k) * NC + r] = wx[k], Mq[(15 + k) * NC + r] = wy[k], Mq[(20 + k) * NC + r] = wp[k];
        // N: row (k,0) has -cos on v_k, row (k,1) -sin on v_k, row (k,2) -tan/wb on v_k and -v sec^2/wb on delta_k
        Mq[(k * 2 + 0) * NC + r] += wx[k] * nd[0] + wy[k] * nd[1] + wp[k] * nd[2];
        Mq[(k * 2 + 1) * NC + r] += wp[k] * nd[3];
        const double c0 = S.cA[0][k + 1] * idt;
        g0 += wx[k] * c0, g1 += wy[k] * c0, g2 += wp[k] * c0;
        const int rd = (k + 1) * 5;
        sgd += wx[k] * gd[rd] + wy[k] * gd[rd + 1] + wp[k] * gd[rd + 2];
        srr += wx[k] * rr[rd] + wy[k] * rr[rd + 1] + wp[k] * rr[rd + 2];
      }
      G0[r * 7 + 0] -= g0, G0[r * 7 + 1] -= g1, G0[r * 7 + 2] -= g2;
      gd[r] -= sgd, rr[r] -= srr;
    }
  }
  OBCA_WARP_SYNC();
  prof_mark(ctx, 16);
  // Householder QR with rank test (LAPACK dgeqr2 reflector convention: v[rk] = 1 implicit)
  int rk = 0, ndrop = 0, nem = 0;
  for (int j = 0; j < nr; ++j) {
    if (j >= 5 && j < 30) continue;  // defining rows of the eliminated variables
    double full = 0, nrm = 0;
#if defined(__CUDA_ARCH__)
    {
      const int lane = ctx.tid & 31;
      double head = 0, tail = 0;
      for (int q = lane; q < NU2; q += 32) {
        double v = Mq[q * NC + j];
        if (q < rk) head += v * v;
        else if (q > rk) tail += v * v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        head += __shfl_xor_sync(0xffffffffu, head, o);
        tail += __shfl_xor_sync(0xffffffffu, tail, o);
      }
      full = head, nrm = tail;
    }
#else
    for (int q = 0; q < NU2; ++q) {
      double v = Mq[q * NC + j];
      if (q < rk) full += v * v;
      else if (q > rk) nrm += v * v;
    }
#endif
    double alpha = rk < NU2 ? Mq[rk * NC + j] : 0.0;
    OBCA_WARP_SYNC();  // every lane holds the pivot entry before lane 0 overwrites it with beta (racecheck: intra-warp read / write hazard)
    double beta = sqrt(alpha * alpha + nrm);
    full = sqrt(full + alpha * alpha + nrm);
    if (rk >= NU2 || !(beta > 1e-8 * fmax(full, rref[j])) || !(full > 0)) {
      // dependent row j = sum_i al[i] * (staircase row i): its remainder is an implied constraint on (xi, dt)
      if (ndrop >= NDR) {
        { *ok = 0; OBCA_DBG("ns fail line %d a=%d i=%d rk=%d ndrop=%d nem=%d nr=%d\n", 663, a, i, rk, ndrop, nem, nr); }
        continue;
      }
      OBCA_LANES(lane) {
        if (lane == 0) {
          double* dr = QRg + QR_DROP + ndrop * DRSZ;
          double* al = misc;
          for (int q = rk - 1; q >= 0; --q) {
            double sacc = Mq[q * NC + j];
            for (int m = q + 1; m < rk; ++m) sacc -= al[m] * Mq[q * NC + (int)piv[m]];
            al[q] = sacc / Mq[q * NC + (int)piv[q]];
          }
          double h[9];
          for (int m = 0; m < 7; ++m) h[m] = G0[j * 7 + m];
          h[7] = gd[j], h[8] = rr[j];
          double scale = fabs(gd[j]);
          for (int m = 0; m < 7; ++m) scale = fmax(scale, fabs(G0[j * 7 + m]));
          for (int q = 0; q < rk; ++q) {
            int jp = (int)piv[q];
            for (int m = 0; m < 7; ++m) h[m] -= al[q] * G0[jp * 7 + m];
            h[7] -= al[q] * gd[jp];
            h[8] -= al[q] * rr[jp];
            dr[3 + q] = al[q];
          }
          double hmax = fabs(h[7]);
          for (int m = 0; m < 7; ++m) hmax = fmax(hmax, fabs(h[m]));
          int slot = -1;
          misc[40] = 0.0;
          if (i > 0 && hmax > 1e-7 * fmax(1.0, scale)) {
            if (nem < NEX) {
              slot = nem;
              misc[40] = 1.0;
              for (int m = 0; m < 9; ++m) em[1 + slot * 9 + m] = h[m];
            } else
              { *ok = 0; OBCA_DBG("ns fail line %d a=%d i=%d rk=%d ndrop=%d nem=%d nr=%d\n", 697, a, i, rk, ndrop, nem, nr); }
          }
          dr[0] = (double)j, dr[1] = (double)slot, dr[2] = (double)rk;
          OBCA_DBG("ns drop a=%d i=%d row=%d rk=%d slot=%d hmax=%.3e scale=%.3e h7=%.3e h8=%.3e beta=%.3e full=%.3e rref=%.3e\n", a, i, j, rk, slot, hmax, scale, h[7], h[8], beta, full, rref[j]);
          for (int m = 0; m < 9; ++m) dr[3 + 35 + m] = h[m];
        }
      }
      OBCA_WARP_SYNC();
      if (misc[40] != 0.0) ++nem;
      OBCA_WARP_SYNC();
      ++ndrop;
      continue;
    }
    if (alpha > 0) beta = -beta;
    const double ibeta = 1.0 / beta;
    const double t = (beta - alpha) * ibeta;
    const double sc = 1.0 / (alpha - beta);
    OBCA_LANES(lane) {
      for (int q = rk + 1 + lane; q < NU2; q += 32) Mq[q * NC + j] *= sc;
      if (lane == 0) tau[rk] = t, piv[rk] = (double)j, Mq[rk * NC + j] = beta, wred[rk] = ibeta;
    }
    OBCA_WARP_SYNC();
    OBCA_LANES(lane) {
      for (int cc = j + 1 + lane; cc < nr; cc += 32) {
        if (cc >= 5 && cc < 30) continue;
        const double* uj = Mq + j;
        double* vc = Mq + cc;
        double a0 = vc[rk * NC], a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int q = rk + 1;
        for (; q + 3 < NU2; q += 4) {
          const double u0 = uj[q * NC], u1 = uj[(q + 1) * NC], u2 = uj[(q + 2) * NC], u3 = uj[(q + 3) * NC];
          const double v0 = vc[q * NC], v1 = vc[(q + 1) * NC], v2 = vc[(q + 2) * NC], v3 = vc[(q + 3) * NC];
          a0 += u0 * v0, a1 += u1 * v1, a2 += u2 * v2, a3 += u3 * v3;
        }
        for (; q < NU2; ++q) a0 += uj[q * NC] * vc[q * NC];
        const double sacc = ((a0 + a1) + (a2 + a3)) * t;
        vc[rk * NC] -= sacc;
        q = rk + 1;
        for (; q + 3 < NU2; q += 4) {  // column cc != column j: load everything first, then store
          const double u0 = uj[q * NC], u1 = uj[(q + 1) * NC], u2 = uj[(q + 2) * NC], u3 = uj[(q + 3) * NC];
          const double v0 = vc[q * NC], v1 = vc[(q + 1) * NC], v2 = vc[(q + 2) * NC], v3 = vc[(q + 3) * NC];
          vc[q * NC] = v0 - sacc * u0, vc[(q + 1) * NC] = v1 - sacc * u1, vc[(q + 2) * NC] = v2 - sacc * u2, vc[(q + 3) * NC] = v3 - sacc * u3;
        }
        for (; q < NU2; ++q) vc[q * NC] -= sacc * uj[q * NC];
      }
    }
    OBCA_WARP_SYNC();
    ++rk;
  }
  int np = NU2 - rk;
  if (np > NP) {
    { *ok = 0; OBCA_DBG("ns fail line %d a=%d i=%d rk=%d ndrop=%d nem=%d nr=%d\n", 747, a, i, rk, ndrop, nem, nr); }
    np = NP;
  }
  prof_mark(ctx, 17);
  // T columns (one lane per column): 0..6 xi, 7..7+NP-1 p, IDT dt; s0
  OBCA_LANES(lane) {
    if (lane == 0) em[0] = (double)nem, QRg[QR_META + 0] = (double)rk, QRg[QR_META + 1] = (double)nr, QRg[QR_META + 2] = (double)ndrop;
    if (lane < 9 + NP) {
      // The column lives in registers (all indices below are compile-time constants after unrolling): the shared-memory
      // pipe, shared by the 8 warps of the CTA, then only serves the broadcast loads of R and of the reflectors.
      const int col = lane;
      double v[NU2];
#pragma unroll
      for (int q = 0; q < NU2; ++q) v[q] = 0.0;
      if (col < 9) {
        // b = -G0[:,col] (col < 7), -gd (col 7), -r (col 8); forward substitution R' w = b on the staircase
#pragma unroll
        for (int ii = 0; ii < NU2; ++ii) {
          if (ii < rk) {
            const int j = (int)piv[ii];
            const double* Rj = Mq + j;
            double b0 = col < 7 ? -G0[j * 7 + col] : (col == 7 ? -gd[j] : -rr[j]), b1 = 0.0;
#pragma unroll
            for (int m = 0; m + 1 < ii; m += 2) b0 -= Rj[m * NC] * v[m], b1 -= Rj[(m + 1) * NC] * v[m + 1];
            if (ii & 1) b0 -= Rj[(ii - 1) * NC] * v[ii - 1];
            v[ii] = (b0 + b1) * wred[ii];
          }
        }
      } else {
        const int jn = col - 9;
#pragma unroll
        for (int q = 0; q < NU2; ++q) v[q] = (jn < np && q == rk + jn) ? 1.0 : 0.0;
      }
      if (col < 9 || col - 9 < np) {
        // v <- H_0 ... H_{rk-1} v ; reflector i = (1, u_{i+1..34}) stored below the staircase of its pivot column
#pragma unroll
        for (int i = NU2 - 1; i >= 0; --i) {
          if (i < rk) {
            const double* u = Mq + (int)piv[i];
            double a0 = v[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int r = i + 1; r < NU2; ++r) {
              const double ur = u[r * NC];
              if (((r - i - 1) & 3) == 0) a0 += ur * v[r];
              else if (((r - i - 1) & 3) == 1) a1 += ur * v[r];
              else if (((r - i - 1) & 3) == 2) a2 += ur * v[r];
              else a3 += ur * v[r];
            }
            const double sacc = ((a0 + a1) + (a2 + a3)) * tau[i];
            v[i] -= sacc;
#pragma unroll
            for (int r = i + 1; r < NU2; ++r) v[r] -= sacc * u[r * NC];
          }
        }
      }
      // (x, y, psi) of the interior nodes from their defining rows: t = rhs - N u, psi = Ci t_psi, x = Ci (t_x - Dx psi), ...
      double tx[5], ty[5], tp[5], ps[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double* nd = ndv + k * 5;
        const int rd = (k + 1) * 5;
        const double c0 = S.cA[0][k + 1] * idt;
        const double uv = v[2 * k], ud = v[2 * k + 1];
        double bx = col == 0 ? -c0 : (col == 7 ? -gd[rd] : (col == 8 ? -rr[rd] : 0.0));
        double by = col == 1 ? -c0 : (col == 7 ? -gd[rd + 1] : (col == 8 ? -rr[rd + 1] : 0.0));
        double bp = col == 2 ? -c0 : (col == 7 ? -gd[rd + 2] : (col == 8 ? -rr[rd + 2] : 0.0));
        tx[k] = bx + nd[0] * uv, ty[k] = by + nd[1] * uv, tp[k] = bp + nd[2] * uv + nd[3] * ud;
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) sacc += S.cAi[j][k] * tp[k];
        ps[j] = sacc * dt;
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double* nd = ndv + k * 5;
        tx[k] -= nd[4] * nd[1] * ps[k], ty[k] += nd[4] * nd[0] * ps[k];
      }
      double* vo = col < 7 ? T + col : (col == 7 ? T + IDT : (col == 8 ? s0 : T + 7 + (col - 9)));
      const int stride = col == 8 ? 1 : NRED;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double sx = 0, sy = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) sx += S.cAi[j][k] * tx[k], sy += S.cAi[j][k] * ty[k];
        vo[(j * 7 + 0) * stride] = sx * dt, vo[(j * 7 + 1) * stride] = sy * dt, vo[(j * 7 + 2) * stride] = ps[j];
        vo[(j * 7 + 3) * stride] = v[2 * j], vo[(j * 7 + 4) * stride] = v[2 * j + 1];
      }
    }
  }
  OBCA_WARP_SYNC();
  // control rows of the stage map: a_k = sum_j cA[j][k]/dt v_j + cA[0][k]/dt xi_v + gd dt + r  (w_k likewise with delta)
  OBCA_LANES(lane) {
    for (int e = lane; e < 10 * (9 + NP); e += 32) {
      const int col = e % (9 + NP), kc = e / (9 + NP), k = 1 + kc / 2, q = 3 + kc % 2, re = k * 5 + q;
      double* vo = col < 7 ? T + col : (col == 7 ? T + IDT : (col == 8 ? s0 : T + 7 + (col - 9)));
      const int stride = col == 8 ? 1 : NRED;
      double sacc = col == q ? S.cA[0][k] * idt : (col == 7 ? gd[re] : (col == 8 ? rr[re] : 0.0));
      for (int j = 1; j < NK; ++j) sacc += S.cA[j][k] * idt * vo[((j - 1) * 7 + q) * stride];
      vo[((k - 1) * 7 + q + 2) * stride] = sacc;
    }
  }
  OBCA_WARP_SYNC();
  prof_mark(ctx, 18);
  // QR record and T map to global memory (multiplier recovery, Riccati dynamics, primal expansion)
  double* Tg = W.TT + (size_t)(a * L.Nmax + i) * (NW * NRED + NW);
  OBCA_LANES(lane) {
    for (int q = lane; q < NQ * NC; q += 32) QRg[q] = Mq[q];
    for (int q = lane; q < NQ; q += 32) QRg[QR_TAU + q] = tau[q], QRg[QR_PIV + q] = piv[q];
    for (int q = lane; q < 2 * NC; q += 32) QRg[QR_BU + q] = bu[q];
    for (int q = lane; q < NW * NRED + NW; q += 32) Tg[q] = T[q];  // T and s0 are contiguous
  }
  OBCA_WARP_SYNC();
  prof_mark(ctx, 19);
  // projected stage Hessian M = Tt' H Tt + dt cross terms, gradient m = Tt'(H s0 + gn) + e_dt hd's0.
  // The node Hessians are expanded to full 7 x 7 blocks first: the inner loops then run over plain strided arrays with
  // compile-time trip counts (the packed-index arithmetic cost more than the multiplications it fed).
  double* HT = Mq;                    // [42][NRED]   (the QR matrix and G0 are no longer needed in shared memory)
  double* hs0 = HT + NS * NRED;       // [42]
  double* Hf = hs0 + NS;              // [6][7][7]
  double* hdT = Hf + NK * 49;         // [NRED] + hds0
  OBCA_LANES(lane) {
    for (int e = lane; e < NK * 49; e += 32) {
      const int k = e / 49, rc = e % 49, r = rc / 7, c = rc % 7;
      Hf[e] = W.HN[((size_t)a * 28 + sym(r, c)) * L.Mv + n0 + k];
    }
  }
  OBCA_WARP_SYNC();
  OBCA_LANES(lane) {
    for (int e = lane; e < NZ * NRED; e += 32) {  // node 0: identity rows of the stage map
      const int row = e / NRED, col = e % NRED;
      HT[e] = col < NZ ? Hf[row * 7 + col] : 0.0;
    }
    // HT rows of the nodes k = 1..K: (7 x 7 node Hessian) x (7 x NRED block of T).  The phase is bound by shared-memory loads (8 warps
    // on one LDS pipe), so the products are register tiled: a lane owns (k, four columns) -- 14 vector loads of T and the 49 Hessian
    // entries (broadcast loads) feed 196 FMAs, 0.3 loads per FMA instead of 2.
    static_assert(NRED == 16 && NZ == 7, "lane mapping written for NRED = 16");
    if (lane < (NK - 1) * 4) {
      const int k = lane / 4 + 1, c0 = (lane % 4) * 4;
      double tc[NZ][4];
#pragma unroll
      for (int m = 0; m < NZ; ++m) {
        const double* tp = T + ((k - 1) * NZ + m) * NRED + c0;
#if defined(__CUDA_ARCH__)
        const double2 v0 = *reinterpret_cast<const double2*>(tp), v1 = *reinterpret_cast<const double2*>(tp + 2);
        tc[m][0] = v0.x, tc[m][1] = v0.y, tc[m][2] = v1.x, tc[m][3] = v1.y;
#else
        for (int j = 0; j < 4; ++j) tc[m][j] = tp[j];
#endif
      }
#pragma unroll
      for (int q = 0; q < NZ; ++q) {
        const double* Hr = Hf + k * 49 + q * 7;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int m = 0; m < NZ; ++m) {
          const double h = Hr[m];
          a0 = fma(h, tc[m][0], a0), a1 = fma(h, tc[m][1], a1), a2 = fma(h, tc[m][2], a2), a3 = fma(h, tc[m][3], a3);
        }
        double* o = HT + (NZ + (k - 1) * NZ + q) * NRED + c0;
        o[0] = a0, o[1] = a1, o[2] = a2, o[3] = a3;
      }
    }
    for (int row = lane; row < NS; row += 32) {
      double acc0 = 0;
      if (row >= NZ) {
        const int k = row / NZ, q = row % NZ;
        const double* Hr = Hf + k * 49 + q * 7;
        const double* sv = s0 + (k - 1) * NZ;
#pragma unroll
        for (int m = 0; m < NZ; ++m) acc0 += Hr[m] * sv[m];
      }
      hs0[row] = acc0;
    }
    if (lane < NRED) {
      double sacc = lane < NZ ? hdv[lane] : 0.0, s1 = 0.0;
      int row = NZ;
      for (; row + 1 < NS; row += 2) sacc += hdv[row] * T[(row - NZ) * NRED + lane], s1 += hdv[row + 1] * T[(row + 1 - NZ) * NRED + lane];
      for (; row < NS; ++row) sacc += hdv[row] * T[(row - NZ) * NRED + lane];
      hdT[lane] = sacc + s1;
    } else if (lane == NRED) {
      double sacc = 0;
      for (int row = NZ; row < NS; ++row) sacc += hdv[row] * s0[row - NZ];
      hdT[NRED] = sacc;
    }
  }
  OBCA_WARP_SYNC();
  double* Mo = W.MA + (size_t)(a * L.Nmax + i) * (NSYM + NRED);
  OBCA_LANES(lane) {
    // M = Tt' HT (lower triangle, packed): a lane owns rows q0, q0 + 1 and columns c0 .. c0 + 3 of the 16 x 16 product -- per row of T one
    // vector load of T and two of HT for 8 FMAs (the tiles above the diagonal are skipped)
    {
      const int q0 = (lane / 4) * 2, c0 = (lane % 4) * 4;
      if (q0 + 1 >= c0) {
        double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        for (int row = 0; row < NW; ++row) {
          const double* tp = T + row * NRED + q0;
          const double* hp = HT + (NZ + row) * NRED + c0;
#if defined(__CUDA_ARCH__)
          const double2 t = *reinterpret_cast<const double2*>(tp), h01 = *reinterpret_cast<const double2*>(hp), h23 = *reinterpret_cast<const double2*>(hp + 2);
          const double tq[2] = {t.x, t.y}, hv[4] = {h01.x, h01.y, h23.x, h23.y};
#else
          const double tq[2] = {tp[0], tp[1]}, hv[4] = {hp[0], hp[1], hp[2], hp[3]};
#endif
#pragma unroll
          for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = fma(tq[a2], hv[b2], acc[a2][b2]);
        }
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int q = q0 + a2, cc = c0 + b2;
            if (cc > q) continue;
            double sacc = acc[a2][b2] + (q < NZ ? HT[q * NRED + cc] : 0.0);  // node-0 identity rows
            if (q == IDT) sacc += hdT[cc];
            if (cc == IDT) sacc += hdT[q];
            Mo[q * (q + 1) / 2 + cc] = sacc;
          }
      }
    }
    if (lane < NRED) {
      const int q = lane;
      double sacc = q < NZ ? hs0[q] + gnv[q] : 0.0;
      for (int row = NZ; row < NS; ++row) sacc += T[(row - NZ) * NRED + q] * (hs0[row] + gnv[row]);
      if (q == IDT) sacc += hdT[NRED];
      Mo[NSYM + q] = sacc;
    }
  }
  OBCA_WARP_SYNC();
  prof_mark(ctx, 20);
}

// All blocks in parallel (one warp each); blocks whose Jacobian is rank deficient hand implied rows to the previous
// interval, which is then re-processed in the next pass (buffers are double-buffered by pass parity so that the passes
// are race-free and deterministic).  Usually two passes: the last interval of a vehicle with an axis-aligned final approach.
OBCA_HDN void interval_nullspace(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, int* ok, int* again, double* arena) {
  assume_scratch(W);
  OBCA_ASSUME_STATIC(L, S);
  const int nblk = L.V * L.Nmax;
#if defined(__CUDA_ARCH__)
  const int wid = ctx.tid >> 5, nw = (ctx.nt >> 5) < OBCA_NS_WARPS ? (ctx.nt >> 5) : OBCA_NS_WARPS;
#else
  const int wid = 0, nw = 1;
#endif
  double* sw = arena + (size_t)(wid < nw ? wid : 0) * NSW;
  for (int it = ctx.tid; it < nblk; it += ctx.nt) {
    W.DF[it] = 1.0, W.DF[nblk + it] = 0.0;
    W.EM[(size_t)it * EXSZ] = 0.0, W.EM[(size_t)(nblk + it) * EXSZ] = 0.0;
  }
  if (ctx.tid == 0) W.EM[(size_t)2 * nblk * EXSZ] = 0.0;
  cta_sync(ctx);
  for (int pass = 0; pass <= L.Nmax; ++pass) {
    const int cur = pass & 1, nxt = cur ^ 1;
    if (ctx.tid == 0) *again = 0;
    for (int it = ctx.tid; it < nblk; it += ctx.nt) W.DF[nxt * nblk + it] = 0.0;
    cta_sync(ctx);
    for (int it = wid; it < nblk && wid < nw; it += nw) {
      int a = it / L.Nmax, i = it % L.Nmax;
      if (i >= L.N[a]) continue;
      const double* em_old = W.EM + (size_t)(cur * nblk + it) * EXSZ;
      double* em_new = W.EM + (size_t)(nxt * nblk + it) * EXSZ;
      if (W.DF[cur * nblk + it] == 0.0) {
        OBCA_LANES(lane) {
          for (int q = lane; q < EXSZ; q += 32) em_new[q] = em_old[q];
        }
        continue;
      }
      // the last block of a vehicle receives no implied rows: point at the always-empty record behind the two buffers
      const double* ex = (i + 1 < L.N[a]) ? W.EM + (size_t)(cur * nblk + it + 1) * EXSZ : W.EM + (size_t)2 * nblk * EXSZ;
      nullspace_block(ctx, L, S, W, a, i, ex, em_new, ok, sw);
      OBCA_LANES(lane) {
        if (lane == 0) {
          bool changed = em_new[0] != em_old[0];
          for (int q = 1; q < 1 + 9 * (int)em_new[0] && !changed; ++q)
            changed = fabs(em_new[q] - em_old[q]) > 1e-12 * fmax(1.0, fabs(em_new[q]));
          if (changed && i > 0) {
            W.DF[nxt * nblk + it - 1] = 1.0;
            *again = 1;
          }
        }
      }
    }
    cta_sync(ctx);
    if (!*again) break;
    cta_sync(ctx);
  }
}

// reduced-coordinate row of stage variable (node k, comp m) of vehicle-interval map T
OBCA_HD double tt_entry(const double* T, int k, int m, int col) {
  if (k == 0) return m == col ? 1.0 : 0.0;
  return T[((k - 1) * NZ + m) * NRED + col];
}

// cross-vehicle coupling: one warp per (pair, interval); lanes own the output entries
//   Mab[ra][cb] = sum_k Ta_k[:,ra]' Hc_k Tb_k[:,cb]  (pose rows of the T maps), plus the two gradient pieces
OBCA_HDN void interval_cross(const Ctx& ctx, const Lay& L, const Scratch& W, double* arena) {
  assume_scratch(W);
#if defined(__CUDA_ARCH__)
  const int wid = ctx.tid >> 5, nw = (ctx.nt >> 5) < OBCA_NS_WARPS ? (ctx.nt >> 5) : OBCA_NS_WARPS;
#else
  const int wid = 0, nw = 1;
#endif
  OBCA_ASSUME_SHARED(arena);
  double* sw = arena + (size_t)(wid < nw ? wid : 0) * NSW;
  // inputs of one (pair, interval) task, double buffered: the next task's operands are fetched from the per-slot work area
  // (global memory, written by the null-space phase) with asynchronous copies while the current task is computed -- under full
  // load (148 CTAs streaming their work areas) the L2/HBM round trip otherwise sits on the critical path of every task
  constexpr int CIN = 2 * NK * 3 * NRED + 2 * NK * 3 + NK * 9;  // ta, tb, sa, sb, hc
  static_assert(2 * CIN + NK * 3 * (NRED + 1) <= NSW, "cross-coupling buffers exceed the warp's shared-memory area");
  double* hb = sw + 2 * CIN;  // [18][NRED + 1] Hc Tb and Hc s0b
  auto fetch = [&](int it, double* in) {
    double* ta = in;                  // [6][3][NRED] pose rows of Ta
    double* tb = ta + NK * 3 * NRED;  // [6][3][NRED]
    double* sa = tb + NK * 3 * NRED;  // [6][3] pose entries of s0a
    double* sb = sa + NK * 3;         // [6][3]
    double* hc = sb + NK * 3;         // [6][9] rows pose_a, cols pose_b
    const int p = it / L.Nmax, i = it % L.Nmax;
    const int a = L.pa[p], b = L.pb[p];
    const double* Ta = W.TT + (size_t)(a * L.Nmax + i) * (NW * NRED + NW);
    const double* Tb = W.TT + (size_t)(b * L.Nmax + i) * (NW * NRED + NW);
    OBCA_LANES(lane) {
      for (int e = lane; e < NK * 3 * NRED; e += 32) {
        const int k = e / (3 * NRED), r = (e / NRED) % 3, col = e % NRED;
        if (k == 0) ta[e] = tb[e] = r == col ? 1.0 : 0.0;  // node 0 is the state itself
        else {
          const int o = ((k - 1) * NZ + r) * NRED + col;
#if defined(__CUDA_ARCH__)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(ta + e)), "l"(Ta + o) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(tb + e)), "l"(Tb + o) : "memory");
#else
          ta[e] = Ta[o], tb[e] = Tb[o];
#endif
        }
      }
      for (int e = lane; e < NK * 3; e += 32) {
        const int k = e / 3, r = e % 3;
        if (k == 0) sa[e] = sb[e] = 0.0;
        else {
          const int o = NW * NRED + (k - 1) * NZ + r;
#if defined(__CUDA_ARCH__)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sa + e)), "l"(Ta + o) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sb + e)), "l"(Tb + o) : "memory");
#else
          sa[e] = Ta[o], sb[e] = Tb[o];
#endif
        }
      }
      for (int e = lane; e < NK * 9; e += 32) {
        const int k = e / 9, r = (e / 3) % 3, m = e % 3;
        const double* src = W.PH + ((size_t)p * 27 + sym(3 + m, r)) * L.Mv + i * NK + k;
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(hc + e)), "l"(src) : "memory");
#else
        hc[e] = *src;
#endif
      }
    }
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  };
  // the tasks of this warp: it = wid, wid + nw, ... over (pair, interval), skipping intervals beyond the pair's horizon
  auto next_task = [&](int it) {
    while (it < L.P * L.Nmax && (it % L.Nmax) * NK >= L.Mp[it / L.Nmax]) it += nw;
    return it;
  };
  int it = wid < nw ? next_task(wid) : L.P * L.Nmax, buf = 0;
  if (it < L.P * L.Nmax) fetch(it, sw);
  for (; it < L.P * L.Nmax;) {
    const int p = it / L.Nmax, i = it % L.Nmax;
    const int itn = next_task(it + nw);
    if (itn < L.P * L.Nmax) fetch(itn, sw + (buf ^ 1) * CIN);
#if defined(__CUDA_ARCH__)
    if (itn < L.P * L.Nmax) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    OBCA_WARP_SYNC();
    const double* ta = sw + buf * CIN;
    const double* tb = ta + NK * 3 * NRED;
    const double* sa = tb + NK * 3 * NRED;
    const double* sb = sa + NK * 3;
    const double* hc = sb + NK * 3;
    double* Mo = W.MAB + (size_t)(p * L.Nmax + i) * (NRED * NRED + 2 * NRED);
    static_assert(NRED == 16, "the lane mapping below is written for NRED = 16");
    // Register-tiled products (the phase is bound by shared-memory loads, 8 warps on one LDS pipe): a lane owns column cb = lane % 16
    // and, with hf = lane / 16, one half of the k's (first product) or eight rows ra (second product).
    //   HB[k][r][cb] = sum_m Hc_k[r][m] Tb_k[m][cb]          -- the three Tb entries of (k, cb) are loaded once for the three rows r
    //   Mab[ra][cb]  = sum_{k,r} Ta_k[r][ra] HB[k][r][cb]    -- one HB entry and eight contiguous Ta entries per (k, r): 5 loads for 8 FMAs
    OBCA_LANES(lane) {
      const int cb = lane & 15, hf = lane >> 4;
      for (int k = hf; k < NK; k += 2) {
        const double t0 = tb[(k * 3 + 0) * NRED + cb], t1 = tb[(k * 3 + 1) * NRED + cb], t2 = tb[(k * 3 + 2) * NRED + cb];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double* h = hc + (k * 3 + r) * 3;
          hb[(k * 3 + r) * (NRED + 1) + cb] = (h[0] * t0 + h[1] * t1) + h[2] * t2;
        }
      }
      if (lane < NK * 3) {  // column NRED: Hc_k s0b
        const int kr = lane, k = kr / 3;
        hb[kr * (NRED + 1) + NRED] = (hc[kr * 3] * sb[k * 3] + hc[kr * 3 + 1] * sb[k * 3 + 1]) + hc[kr * 3 + 2] * sb[k * 3 + 2];
      }
    }
    OBCA_WARP_SYNC();
    OBCA_LANES(lane) {
      const int cb = lane & 15, hf = lane >> 4;
      double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
      for (int kr = 0; kr < NK * 3; ++kr) {
        const double h = hb[kr * (NRED + 1) + cb];
        const double* t = ta + kr * NRED + hf * 8;  // 16-byte aligned: vectorised loads
#if defined(__CUDA_ARCH__)
        const double2 a01 = *reinterpret_cast<const double2*>(t), a23 = *reinterpret_cast<const double2*>(t + 2);
        const double2 a45 = *reinterpret_cast<const double2*>(t + 4), a67 = *reinterpret_cast<const double2*>(t + 6);
        acc[0] = fma(a01.x, h, acc[0]), acc[1] = fma(a01.y, h, acc[1]), acc[2] = fma(a23.x, h, acc[2]), acc[3] = fma(a23.y, h, acc[3]);
        acc[4] = fma(a45.x, h, acc[4]), acc[5] = fma(a45.y, h, acc[5]), acc[6] = fma(a67.x, h, acc[6]), acc[7] = fma(a67.y, h, acc[7]);
#else
        for (int j = 0; j < 8; ++j) acc[j] = fma(t[j], h, acc[j]);
#endif
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Mo[(hf * 8 + j) * NRED + cb] = acc[j];
      // the two gradient pieces: Ta' Hc s0b (lanes 0..15) and Tb' Hc' s0a (lanes 16..31)
      double g = 0;
      if (hf == 0) {
        for (int kr = 0; kr < NK * 3; ++kr) g = fma(ta[kr * NRED + cb], hb[kr * (NRED + 1) + NRED], g);
      } else {
        for (int k = 0; k < NK; ++k)
          for (int m = 0; m < 3; ++m) {
            double h = 0;
            for (int r = 0; r < 3; ++r) h += hc[k * 9 + r * 3 + m] * sa[k * 3 + r];
            g += tb[(k * 3 + m) * NRED + cb] * h;
          }
      }
      Mo[NRED * NRED + lane] = g;
    }
    OBCA_WARP_SYNC();
    it = itn, buf ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------
// [RICCATI]  stage i: state X = (xi_1 .. xi_V, dt), control U = free directions of the active vehicles (compacted,
// nu_i = sum_a np_a(i) <= NP V).  The dynamics are block diagonal per vehicle plus a dt column:
//   xi_a' = Aa xi_a + da dt + Ba u_a + ca   (rows 28..34 of the block's T map),   dt' = dt.
// All stage matrices live in the work arena RW (shared memory on the device).
// ------------------------------------------------------------------------------------------------
// pivot row / column exchange buffers of the stage factorisation (cta_stage_ldl)
struct LdlBuf {
  double col[2][64], row[2][128], inv[2], invd[64];
  int bad;
};
#if defined(__CUDA_ARCH__)
#define OBCA_EMU_LDL 0
#else
#define OBCA_EMU_LDL 1  // the host emulation factorises in the same L D L' form (stage_ldl_scalar)
#endif

// doubles of one stage record of the forward pass: Ks' [nX + 1][nU] and L [nU][nU]
OBCA_HD size_t ric_record_doubles(const Lay& L) { return (size_t)(L.nX + 1) * L.nU + (size_t)L.nU * L.nU; }

struct RicWork {
  double *P, *p, *Q, *S, *R, *q, *r, *PA, *PB, *pc, *F, *Gm, *K, *Ab, *Bb, *db, *cb, *invd, *MAs, *MABs, *TBs;
  int *uoff, *npv, *npt;  // [MAXV + 1], [MAXV + 1], [V][Nmax]
  const double** isrc;     // [n_in] address of every stage input at stage 0
  int* istr;               // [n_in][2] (doubles per stage, first stage beyond the block's horizon)
  int* rmap;               // [7 V + nUmax] row / column of the stage matrix [states; controls] -> vehicle * 32 + reduced coordinate (-1: unused)
  int* btab;               // [V][V] block (a, b) of the stage Hessian: 4 * (offset from MAs) + kind (0 zero, 1 packed diagonal block, 2 MAB, 3 MAB transposed)
};

inline size_t riccati_only_doubles(const Lay& L) {
  size_t nX = L.nX, nU = L.nU;
  return 4 * nX * nX + 4 * nX * nU + 2 * nU * nU + 2 * nU * (nX + 1) + 8 * (nX + nU) + (size_t)L.V * (49 + 7 * NP + 14) + 2 * (MAXV + 1) + 32 + nU + ((size_t)L.V * L.Nmax + 2) / 2 +
         (OBCA_RIC_PREFETCH ? 3 * ((size_t)L.V * (NSYM + NRED) + (size_t)L.P * (NRED * NRED + 2 * NRED) + (size_t)L.V * 7 * (NRED + 1)) + 4 : 0) +
         (nX + nU) / 2 + 2 + (size_t)(L.V * L.V + 1) / 2;
}

// work arena shared by the null-space phase (NSW doubles per warp) and the Riccati phase
inline size_t riccati_work_doubles(const Lay& L, int nwarps) {
  size_t a = riccati_only_doubles(L), b = (size_t)nwarps * NSW;
  return a > b ? a : b;
}

OBCA_HD int ric_input_count(const Lay& L);
// SM: the arena is shared memory (V <= 4: everything of a stage fits); otherwise it is the per-slot global-memory arena
// (the state of V > 4 vehicles, 7 V + 1 > 29, no longer fits next to the null-space work areas) -- same code, other address space
template <bool SM>
OBCA_HD void ric_carve(RicWork& R, const Lay& L, double* w) {
  if (SM) OBCA_ASSUME_SHARED(w);
  int nX = L.nX, nU = L.nU;
  R.P = w, w += nX * nX;
  R.p = w, w += nX;
  R.Q = w, w += nX * nX;
  R.S = w, w += nU * nX;
  R.R = w, w += nU * nU;
  R.q = w, w += nX;
  R.r = w, w += nU;
  R.PA = w, w += nX * nX;
  R.PB = w, w += nX * nU;
  R.pc = w, w += nX;
  R.F = w, w += nU * nU;
  R.Gm = w, w += nU * (nX + 1);
  R.K = w, w += nU * (nX + 1);
  R.Ab = w, w += L.V * 49;
  R.Bb = w, w += L.V * 7 * NP;
  R.db = w, w += L.V * 7;
  R.cb = w, w += L.V * 7;
  R.invd = w, w += nU;
  R.MAs = w;
  R.MABs = w + L.V * (NSYM + NRED);
  R.TBs = R.MABs + L.P * (NRED * NRED + 2 * NRED);
  if (OBCA_RIC_PREFETCH) w += L.V * (NSYM + NRED) + L.P * (NRED * NRED + 2 * NRED) + L.V * 7 * (NRED + 1);
  R.uoff = (int*)w;
  R.npv = R.uoff + (MAXV + 1);
  R.npt = R.npv + (MAXV + 1);  // [V][Nmax] free directions of every block
  {
    size_t off = (size_t)((char*)(R.npt + L.V * L.Nmax) - (char*)0);
    off = (off + 15) & ~(size_t)15;
    R.isrc = (const double**)((char*)0 + off);
    R.istr = (int*)(R.isrc + ric_input_count(L));
    R.rmap = R.istr + (OBCA_RIC_PREFETCH ? 2 * ric_input_count(L) : 0);
    R.btab = R.rmap + (7 * L.V + L.nU);
  }
}

// Stage inputs of the Riccati recursion (projected Hessians MA, coupling blocks MAB, last rows of the T maps).  The copies
// for stage i-1 are issued as asynchronous global -> shared copies (cp.async, 8 bytes each, zero fill for blocks beyond a
// vehicle's horizon) once stage i has consumed its inputs, so the global-memory latency hides behind the factorisation.
OBCA_HD int ric_input_count(const Lay& L) { return L.V * (NSYM + NRED) + L.P * (NRED * NRED + 2 * NRED) + L.V * 7 * (NRED + 1); }
OBCA_HD const double* ric_input_addr(const Lay& L, const Scratch& W, int i, int it, int* lim) {
  const int n1 = L.V * (NSYM + NRED), n2 = n1 + L.P * (NRED * NRED + 2 * NRED);
  if (it < n1) {
    const int a = it / (NSYM + NRED);
    *lim = L.N[a];
    return W.MA + (size_t)(a * L.Nmax + i) * (NSYM + NRED) + it % (NSYM + NRED);
  }
  if (it < n2) {
    const int e = it - n1, p = e / (NRED * NRED + 2 * NRED);
    *lim = (L.Mp[p] + NK - 1) / NK;  // active while i * NK < Mp
    return W.MAB + (size_t)(p * L.Nmax + i) * (NRED * NRED + 2 * NRED) + e % (NRED * NRED + 2 * NRED);
  }
  const int e = it - n2, a = e / (7 * (NRED + 1)), r = (e / (NRED + 1)) % 7, cc = e % (NRED + 1);
  *lim = L.N[a];
  const double* T = W.TT + (size_t)(a * L.Nmax + i) * (NW * NRED + NW);
  return cc < NRED ? T + (28 + r) * NRED + cc : T + NW * NRED + 28 + r;
}
// table of the stage inputs (address at stage 0, stride per stage, horizon): the per-stage fetch then needs no divisions
OBCA_HD void ric_input_table(const Ctx& ctx, const Lay& L, const Scratch& W, const RicWork& R) {
  const int tot = ric_input_count(L);
  for (int it = ctx.tid; it < tot; it += ctx.nt) {
    int lim, lim1;
    const double* s0 = ric_input_addr(L, W, 0, it, &lim);
    const double* s1 = ric_input_addr(L, W, 1, it, &lim1);
    R.isrc[it] = s0;
    R.istr[2 * it] = (int)(s1 - s0), R.istr[2 * it + 1] = lim;
  }
}
template <bool SM>
OBCA_HD void ric_input_fetch(const Ctx& ctx, const Lay& L, const Scratch& W, const RicWork& R, int i) {
  const int tot = ric_input_count(L);
  for (int it = ctx.tid; it < tot; it += ctx.nt) {
    const bool active = i < R.istr[2 * it + 1];
    const double* src = R.isrc[it] + (size_t)i * R.istr[2 * it];
#if defined(__CUDA_ARCH__)
    if (SM) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(R.MAs + it)), "l"(active ? src : W.MA),
                   "r"(active ? 8 : 0)
                   : "memory");
      continue;
    }
#endif
    R.MAs[it] = active ? *src : 0.0;  // MAs, MABs, TBs are contiguous
  }
}
OBCA_HD void ric_input_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// free directions of block (a, i): 10 - rank, 0 for a vehicle whose horizon has ended
OBCA_HD int block_np(const Lay& L, const Scratch& W, int a, int i) {
  if (i >= L.N[a]) return 0;
  int np = NU2 - (int)W.QR[(size_t)(a * L.Nmax + i) * QRSZ + QR_META];
  return np > NP ? NP : np;
}

// reduced coordinate of vehicle a -> (kind, index): kind 0 = state index, 1 = control index, -1 = unused control slot
OBCA_HD int red_target(int a, int rc, int V, const int* uoff, const int* npv, int* kind) {
  if (rc < 7) {
    *kind = 0;
    return 7 * a + rc;
  }
  if (rc == IDT) {
    *kind = 0;
    return 7 * V;
  }
  if (rc - 7 < npv[a]) {
    *kind = 1;
    return uoff[a] + rc - 7;
  }
  *kind = -1;
  return 0;
}

// Stage assembly as ONE gather pass: every entry of Q, S, R (and of the dt row / column and the gradients q, r) is computed by
// one thread from the prefetched stage inputs -- no zero fill, no scatter passes, one barrier (round 1: zero fill + three scatter
// passes with three barriers, 11 k cycles per stage).
//   state index t < 7V -> (vehicle t / 7, reduced coordinate t % 7); control u -> (vehicle, 7 + slot) through ctl[]; dt = 7V
//   own-vehicle entries come from MA (packed symmetric 16 x 16), cross-vehicle entries from MAB of the pair (rows: first vehicle)
OBCA_HD void riccati_stage_assemble(const Ctx& ctx, const Lay& L, const Scratch& W, const RicWork& R, int i, double hdtdt, int nu) {
  assume_scratch(W);
  const int nX = L.nX, idt = 7 * L.V, V = L.V;
#if OBCA_RIC_PREFETCH
#define OBCA_MA(a) (R.MAs + (a) * (NSYM + NRED))
#define OBCA_MAB(p) (R.MABs + (p) * (NRED * NRED + 2 * NRED))
#else
#define OBCA_MA(a) (W.MA + (size_t)((a) * L.Nmax + i) * (NSYM + NRED))
#define OBCA_MAB(p) (W.MAB + (size_t)((p) * L.Nmax + i) * (NRED * NRED + 2 * NRED))
#endif
  prof_mark(ctx, 27);
  // block dynamics from the T maps
  for (int it = ctx.tid; it < V * 7 * (NRED + 1); it += ctx.nt) {
    int a = it / (7 * (NRED + 1)), r = (it / (NRED + 1)) % 7, cc = it % (NRED + 1);
#if OBCA_RIC_PREFETCH
    const double v = R.TBs[it];
#else
    double v = 0.0;
    if (i < L.N[a]) {
      const double* T = W.TT + (size_t)(a * L.Nmax + i) * (NW * NRED + NW);
      v = cc < NRED ? T[(28 + r) * NRED + cc] : T[NW * NRED + 28 + r];
    }
#endif
    if (cc < 7) R.Ab[a * 49 + r * 7 + cc] = v;
    else if (cc < IDT) R.Bb[(a * 7 + r) * NP + cc - 7] = v;
    else if (cc == IDT) R.db[a * 7 + r] = v;
    else R.cb[a * 7 + r] = v;
  }
  prof_mark(ctx, 28);
  // entry (vehicle a, reduced coordinate ra) x (vehicle b, reduced coordinate rb) of the stage Hessian
  auto entry = [&](int a, int ra, int b, int rb) -> double {
#if OBCA_RIC_PREFETCH
    // block table of this stage (filled with the row map): no horizon tests, no pair-index arithmetic per entry
    const int code = R.btab[a * V + b], kind = code & 3;
    const double* base = R.MAs + (code >> 2);
    if (kind == 0) return 0.0;
    return kind == 1 ? base[sym(ra, rb)] : (kind == 2 ? base[ra * NRED + rb] : base[rb * NRED + ra]);
#else
    if (a == b) return i < L.N[a] ? OBCA_MA(a)[sym(ra, rb)] : 0.0;
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    const int p = lo * V - lo * (lo + 1) / 2 + (hi - lo - 1);  // index of the pair (lo, hi) in combinations order
    if (i * NK >= L.Mp[p]) return 0.0;
    return a < b ? OBCA_MAB(p)[ra * NRED + rb] : OBCA_MAB(p)[rb * NRED + ra];
#endif
  };
  // [Q S'; S R]: `tpr` threads per row, each walks a contiguous run of columns (row / column coordinates from R.rmap)
  const int nrow = idt + nu, nD = nrow + 1;
  const int sh = ctx.nt >= 4 * nrow ? 2 : (ctx.nt >= 2 * nrow ? 1 : 0), tpr = 1 << sh;
  for (int t = ctx.tid; t < nrow * tpr; t += ctx.nt) {
    const int row = t >> sh, part = t & (tpr - 1);
    const int cr = R.rmap[row], a = cr >> 5, ra = cr & 31;
    const int ncol = row < idt ? idt : nrow;  // state rows: Q only (S' is not stored)
    const int c0 = (part * ncol) >> sh, c1 = ((part + 1) * ncol) >> sh;
    double* ds = row < idt ? R.Q + row * nX : R.S + (row - idt) * nX;
    double* dc = R.R + (row - idt) * nu - idt;
    for (int c = c0; c < c1; ++c) {
      const int cc = R.rmap[c];
      const double v = entry(a, ra, cc >> 5, cc & 31);
      if (c < idt) ds[c] = v;
      else dc[c] = v;
    }
  }
  // everything that touches dt, and the gradients: one thread per target (the last threads of the CTA), fixed summation order
  for (int t = ctx.nt - 1 - ctx.tid; t < nD; t += ctx.nt) {
    {
      int a = -1, rc = IDT;
      if (t < nrow) a = R.rmap[t] >> 5, rc = R.rmap[t] & 31;
      double hd = 0, g = 0;
      if (a >= 0) {
        if (i < L.N[a]) {
          const double* Mo = OBCA_MA(a);
          hd = Mo[sym(IDT, rc)], g = Mo[NSYM + rc];
        }
        for (int p = 0; p < L.P; ++p) {
          if (i * NK >= L.Mp[p]) continue;
          const double* Mo = OBCA_MAB(p);
          if (L.pa[p] == a) hd += Mo[rc * NRED + IDT], g += Mo[NRED * NRED + rc];
          else if (L.pb[p] == a) hd += Mo[IDT * NRED + rc], g += Mo[NRED * NRED + NRED + rc];
        }
        if (t < idt) R.Q[t * nX + idt] = hd, R.Q[idt * nX + t] = hd, R.q[t] = g;
        else R.S[(t - idt) * nX + idt] = hd, R.r[t - idt] = g;
      } else {
        for (int aa = 0; aa < V; ++aa) {
          if (i >= L.N[aa]) continue;
          const double* Mo = OBCA_MA(aa);
          hd += Mo[sym(IDT, IDT)], g += Mo[NSYM + IDT];
        }
        for (int p = 0; p < L.P; ++p) {
          if (i * NK >= L.Mp[p]) continue;
          const double* Mo = OBCA_MAB(p);
          hd += 2.0 * Mo[IDT * NRED + IDT], g += Mo[NRED * NRED + IDT] + Mo[NRED * NRED + NRED + IDT];
        }
        if (i == 0) hd += hdtdt, g += W.gphi[L.oDT];
        R.Q[idt * nX + idt] = hd, R.q[idt] = g;
      }
    }
  }
  cta_sync(ctx);
}

// ------------------------------------------------------------------------------------------------
// Factorisation of the stage matrix.  F = R + B'PB (nu x nu, nu <= 32) and the right-hand sides Gm = [S + B'PA | r + B'pc]
// (nu x nc) are eliminated together:  F = L D L' (unit L),  Khat = L^-1 Gm.  The cost-to-go update only needs Khat and D
// (Gm' F^-1 Gm = Khat' D^-1 Khat), the controls follow in the forward pass from  L' u = -D^-1 Khat [x; 1].
//
// Device (256 threads = 4 row groups x 64 columns): every thread owns one column of the augmented matrix [F | Gm] and the
// rows rg, rg + 4, ... of it, in registers.  The entries are computed in place (the products B'PB, B'PA, B'pc are 7-term
// dot products), then the right-looking elimination runs with ONE CTA barrier per pivot: the owners of the next pivot row
// and column publish them through double-buffered shared memory, the owner of the pivot publishes its reciprocal
// (hardware seed + two Newton steps).  Round 1 eliminated in shared memory with two barriers per pivot plus a barrier per
// row of the back-substitution: 1.33 M of 8.1 M cycles per iteration; the micro-benchmark of this kernel
// (tools/microbench/ldl_bench.cu) runs a 20 x 50 stage in 7.9 k cycles.
// Outputs (shared memory): F <- unit L below the diagonal (row major, ld nu), Gm <- Khat, Ks <- -D^-1 Khat, invd <- 1 / d.
// A pivot below PIVOT_TOL (relative to R_jj) means the reduced Hessian is not positive definite: *ok = 0.
// ------------------------------------------------------------------------------------------------
OBCA_HD bool ldl_fast_shape(const Ctx& ctx, int nu, int nc) {
#if defined(__CUDA_ARCH__)
  return ctx.nt == 256 && ctx.ldl && nu >= 1 && nu <= 64 && nu + nc <= 128;
#else
  (void)ctx, (void)nu, (void)nc;
  return false;
#endif
}

#if defined(__CUDA_ARCH__)
// LOG2CG: log2 of the columns per row group (64 columns x 4 row groups for up to 4 vehicles, 128 x 2 beyond);
// KMAX: rows per thread (row r = rg + RG k)
template <int LOG2CG, int KMAX>
__device__ __forceinline__ void cta_stage_ldl(const RicWork& R, int nu, int nX, int* ok, LdlBuf* B) {
  constexpr int CG = 1 << LOG2CG, RG = 256 >> LOG2CG;
  OBCA_ASSUME_SHARED(B);  // the exchange buffers are static shared memory of k_solve (LDS / STS instead of generic accesses)
  const int t = threadIdx.x, rg = t >> LOG2CG, c = t & (CG - 1), nc = nX + 1, ncols = nu + nc;
  double m[KMAX];
  // entries of [F | Gm] owned by this thread: F = R + B'PB ; Gm = [S + B'PA | r + B'pc]
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int r = rg + RG * k;
    double v = 0.0;
    if (r < nu && c < ncols) {
      int a = 0;
      while (r >= R.uoff[a + 1]) ++a;
      const double* Bj = R.Bb + (a * 7) * NP + (r - R.uoff[a]);
      const double* rhs;
      int rs;
      if (c < nu) rhs = R.PB + (7 * a) * nu + c, rs = nu, v = R.R[r * nu + c];
      else if (c < nu + nX) rhs = R.PA + (7 * a) * nX + (c - nu), rs = nX, v = R.S[r * nX + (c - nu)];
      else rhs = R.pc + 7 * a, rs = 1, v = R.r[r];
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q += 2) s0 += Bj[q * NP] * rhs[q * rs], s1 += Bj[(q + 1) * NP] * rhs[(q + 1) * rs];
      v += (s0 + Bj[6 * NP] * rhs[6 * rs]) + s1;
    }
    m[k] = v;
  }
  // every column has one diagonal owner, thread (rg = c mod RG, c): its pivot tolerance is known before the loop
  const double mytol = (c < nu && rg == (c & (RG - 1))) ? PIVOT_TOL * fmax(1.0, fabs(R.R[c * nu + c])) : 0.0;
  if (t == 0) B->bad = 0;
  __syncthreads();
  if (c == 0) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) B->col[0][rg + RG * k] = m[k];
  }
  if (rg == 0) B->row[0][c] = m[0];
  if (t == 0) {
    double d = m[0];
    if (!(d > mytol)) B->bad = 1, d = 1.0;
    const double inv = rcp_pos(d);
    B->inv[0] = inv, B->invd[0] = inv;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RG * KMAX; ++j) {
    if (j < nu) {  // CTA-uniform
      const int p = j & 1;
      const double uc = c > j ? B->row[p][c] * B->inv[p] : 0.0;  // finished columns (c <= j) keep their entries: L (times d)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (RG * k + (RG - 1) > j) {  // static: some row of this slot lies below the pivot
          const int r = rg + RG * k;
          const double l = r > j ? B->col[p][r] : 0.0;
          m[k] = fma(-l, uc, m[k]);
        }
      }
      if (j + 1 < nu) {
        const int jn = j + 1, kn = jn / RG, gn = jn & (RG - 1);
        if (c == jn) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (RG * k + (RG - 1) > jn) B->col[p ^ 1][rg + RG * k] = m[k];
        }
        if (rg == gn) {
          B->row[p ^ 1][c] = m[kn < KMAX ? kn : 0];
          if (c == jn) {
            double d = m[kn < KMAX ? kn : 0];
            if (!(d > mytol)) B->bad = 1, d = 1.0;
            const double inv = rcp_pos(d);
            B->inv[p ^ 1] = inv, B->invd[jn] = inv;
          }
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int r = rg + RG * k;
    if (r < nu && c < ncols) {
      if (c < nu) {
        if (r > c) R.F[r * nu + c] = m[k] * B->invd[c];
      } else {
        R.Gm[r * nc + (c - nu)] = m[k];
        R.K[r * nc + (c - nu)] = -m[k] * B->invd[r];
      }
    }
  }
  if (t < nu) R.invd[t] = B->invd[t];
  if (t == 0 && B->bad) *ok = 0;
}
#endif

// fused products + factorisation of one stage (device fast path); all threads of the CTA call it
OBCA_HD void stage_ldl_fused(const Ctx& ctx, const RicWork& R, int nu, int nX, int* ok, LdlBuf* B) {
#if defined(__CUDA_ARCH__)
  (void)ctx;
  if (nu + nX + 1 <= 64) {
    if (nu <= 8) cta_stage_ldl<6, 2>(R, nu, nX, ok, B);
    else if (nu <= 16) cta_stage_ldl<6, 4>(R, nu, nX, ok, B);
    else if (nu <= 24) cta_stage_ldl<6, 6>(R, nu, nX, ok, B);
    else cta_stage_ldl<6, 8>(R, nu, nX, ok, B);
  } else {
    if (nu <= 32) cta_stage_ldl<7, 16>(R, nu, nX, ok, B);
    else if (nu <= 48) cta_stage_ldl<7, 24>(R, nu, nX, ok, B);
    else cta_stage_ldl<7, 32>(R, nu, nX, ok, B);
  }
#else
  (void)ctx, (void)R, (void)nu, (void)nX, (void)ok, (void)B;
#endif
}

// host emulation / reference form of the same factorisation on the assembled F, Gm (one thread)
OBCA_HD void stage_ldl_scalar(double* F, int nu, double* Gm, int nc, double* Ks, double* invd, const double* Rm, int* ok) {
  for (int j = 0; j < nu; ++j) {
    double d = F[j * nu + j];
    if (!(d > PIVOT_TOL * fmax(1.0, fabs(Rm[j * nu + j])))) *ok = 0, d = 1.0;
    const double inv = 1.0 / d;
    invd[j] = inv;
    for (int r = j + 1; r < nu; ++r) {
      const double lr = F[r * nu + j];
      for (int c = j + 1; c < nu; ++c) F[r * nu + c] -= lr * (F[j * nu + c] * inv);
      for (int c = 0; c < nc; ++c) Gm[r * nc + c] -= lr * (Gm[j * nc + c] * inv);
    }
    for (int r = j + 1; r < nu; ++r) F[r * nu + j] *= inv;
  }
  for (int r = 0; r < nu; ++r)
    for (int c = 0; c < nc; ++c) Ks[r * nc + c] = -Gm[r * nc + c] * invd[r];
}

// Stage shapes the register kernel does not cover (more than 32 free directions, V > 4): right-looking Cholesky of the
// augmented matrix [F | Gm] in shared memory by the whole CTA (two barriers per pivot), then a column-oriented backward
// substitution (one barrier per row); leaves the gains K = -F^-1 Gm in R.K and Gm untouched.
OBCA_HDN void riccati_factor_generic(const Ctx& ctx, const RicWork& R, int nu, int nc, int i, int* ok) {
  (void)i;
    // Cholesky F = L L' fused with the forward substitution of [K | k] = -F^-1 Gm: right-looking elimination of the
    // augmented matrix [F | Gm] by the whole CTA (two barriers per pivot); a non-positive pivot means the reduced
    // Hessian is not positive definite.  Then a column-oriented backward substitution (one barrier per row).
    for (int it = ctx.tid; it < nu * nc; it += ctx.nt) R.K[it] = R.Gm[it];
    cta_sync(ctx);
    // thread -> (row group rg, column c): c < 64 covers the n1 trailing columns of F and the nc columns of K without any
    // integer division in the loops; every thread keeps its column
    const int c64 = ctx.tid & 63, rg = ctx.tid >> 6, nrg = ctx.nt >> 6 > 0 ? ctx.nt >> 6 : 1;
    for (int j = 0; j < nu; ++j) {
      double d = R.F[j * nu + j];
      const bool bad = !(d > PIVOT_TOL * fmax(1.0, fabs(R.R[j * nu + j])));
      if (bad) OBCA_DBG("riccati bad pivot stage %d j=%d nu=%d d=%.3e R=%.3e\n", i, j, nu, R.F[j * nu + j], R.R[j * nu + j]);
      if (bad) d = 1.0;
#if defined(__CUDA_ARCH__)
      const double inv = rsqrt(d);
#else
      const double inv = 1.0 / sqrt(d);
#endif
      const int n1 = nu - j - 1;
      for (int it = ctx.tid; it < n1 + nc + 1; it += ctx.nt) {  // the pivot entry itself is left untouched (only 1/L_jj is used later)
        if (it < n1) R.F[(j + 1 + it) * nu + j] *= inv;
        else if (it < n1 + nc) R.K[j * nc + (it - n1)] *= inv;
        else {
          R.invd[j] = inv;
          if (bad) *ok = 0;
        }
      }
      cta_sync(ctx);
#if defined(__CUDA_ARCH__)
      if (n1 + nc <= 64) {
        if (c64 < n1) {
          const int cc = j + 1 + c64;
          const double lcj = R.F[cc * nu + j];
          for (int r = cc + rg; r < nu; r += nrg) R.F[r * nu + cc] -= R.F[r * nu + j] * lcj;
        } else if (c64 < n1 + nc) {
          const int c = c64 - n1;
          const double kj = R.K[j * nc + c];
          for (int r = j + 1 + rg; r < nu; r += nrg) R.K[r * nc + c] -= R.F[r * nu + j] * kj;
        }
      } else
#endif
      {
        for (int it = ctx.tid; it < n1 * (n1 + nc); it += ctx.nt) {
          const int r = j + 1 + it / (n1 + nc), c = it % (n1 + nc);
          const double lrj = R.F[r * nu + j];
          if (c < n1) {
            const int cc = j + 1 + c;
            if (cc <= r) R.F[r * nu + cc] -= lrj * R.F[cc * nu + j];
          } else
            R.K[r * nc + (c - n1)] -= lrj * R.K[j * nc + (c - n1)];
        }
      }
      cta_sync(ctx);
    }
    prof_mark(ctx, 14);
    // backward: rows r = nu-1 .. 0, w_r = K[r,:] stays unscaled until the end; K[m,:] -= L[r][m] invd[r] w_r for m < r
    for (int r = nu - 1; r > 0; --r) {
      const double ir = R.invd[r];
      for (int it = ctx.tid; it < r * nc; it += ctx.nt) {
        const int m = it / nc, c = it % nc;
        R.K[m * nc + c] -= R.F[r * nu + m] * ir * R.K[r * nc + c];
      }
      cta_sync(ctx);
    }
    for (int it = ctx.tid; it < nu * nc; it += ctx.nt) R.K[it] = -R.K[it] * R.invd[it / nc];
    cta_sync(ctx);
    prof_mark(ctx, 15);
}

template <bool SM>
OBCA_HDN void riccati_backward(const Ctx& ctx, const Lay& L, const Scratch& W, double* RW, double hdtdt, int* ok) {
  assume_scratch(W);
  const int nX = L.nX, nUmax = L.nU, idt = 7 * L.V, V = L.V;
  RicWork R;
  ric_carve<SM>(R, L, RW);
  const size_t pstride = (size_t)nX * nX + nX, kstride = ric_record_doubles(L);
  for (int q = ctx.tid; q < nX * nX; q += ctx.nt) R.P[q] = 0;
  for (int q = ctx.tid; q < nX; q += ctx.nt) R.p[q] = 0;
  {
    double* Pn = W.RP + (size_t)L.Nmax * pstride;
    for (int q = ctx.tid; q < (int)pstride; q += ctx.nt) Pn[q] = 0;
  }
  for (int it = ctx.tid; it < V * L.Nmax; it += ctx.nt) R.npt[it] = block_np(L, W, it / L.Nmax, it % L.Nmax);
#if OBCA_RIC_PREFETCH
  ric_input_table(ctx, L, W, R);
  cta_sync(ctx);
  ric_input_fetch<SM>(ctx, L, W, R, L.Nmax - 1);
#endif
  cta_sync(ctx);
  for (int i = L.Nmax - 1; i >= 0; --i) {
    if (ctx.tid == 0) {
      int off = 0;
      for (int a = 0; a < V; ++a) {
        R.uoff[a] = off;
        R.npv[a] = R.npt[a * L.Nmax + i];
        off += R.npv[a];
      }
      R.uoff[V] = off;
    }
    // (vehicle, reduced coordinate) of every row / column of this stage's matrix, once per stage: the gather below then needs no
    // integer divisions by run-time values and no searches
    for (int t = ctx.nt - 1 - ctx.tid; t < idt + nUmax; t += ctx.nt) {
      int code = -1;
      if (t < idt) code = (t / 7) * 32 + t % 7;
      else {
        int u = t - idt, off = 0;
        for (int a = 0; a < V && code < 0; ++a) {
          const int np = R.npt[a * L.Nmax + i];
          if (u < off + np) code = a * 32 + 7 + (u - off);
          off += np;
        }
      }
      R.rmap[t] = code;
    }
#if OBCA_RIC_PREFETCH
    for (int t = ctx.tid; t < V * V; t += ctx.nt) {
      const int a = t / V, b = t % V;
      int code = 0;
      if (a == b) {
        if (i < L.N[a]) code = 4 * (a * (NSYM + NRED)) + 1;
      } else {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int p = lo * V - lo * (lo + 1) / 2 + (hi - lo - 1);  // index of the pair (lo, hi) in combinations order
        if (i * NK < L.Mp[p]) code = 4 * (V * (NSYM + NRED) + p * (NRED * NRED + 2 * NRED)) + (a < b ? 2 : 3);
      }
      R.btab[t] = code;
    }
#endif
#if OBCA_RIC_PREFETCH
    ric_input_wait();  // this stage's inputs (issued during the previous stage) have landed; the barrier publishes them
#endif
    cta_sync(ctx);
    const int nu = R.uoff[V];
    prof_mark(ctx, 6);
    riccati_stage_assemble(ctx, L, W, R, i, hdtdt, nu);
    prof_mark(ctx, 12);
    // PA = P A ; PB = P B ; pc = P c + p   (block structure of A, B).  The 7-term items and the two dense columns (dt and the
    // constant) are separate loops so that the lanes of a warp do the same amount of work; the dense ones go to the last threads.
    {
      // a thread owns one column (64 or 128 columns per pass) and every nrg-th row: the 7 coefficients of the column stay in registers
      // and no index needs a division by a run-time value; a single thread (host emulation) walks all columns
      const int ncols = idt + nu;
      const bool grid2d = ctx.nt >= 128 && (ctx.nt & (ctx.nt - 1)) == 0;
      const int lcg = ncols <= 64 ? 6 : 7;
      const int cstep = grid2d ? (1 << lcg) : 1, rg = grid2d ? ctx.tid >> lcg : 0, nrg = grid2d ? ctx.nt >> lcg : 1;
      for (int col = grid2d ? (ctx.tid & (cstep - 1)) : 0; col < ncols; col += grid2d ? ncols : 1) {
        const double* src;
        double* dst;
        int a, sstr, ld;
        if (col < idt) a = col / 7, src = R.Ab + a * 49 + col % 7, sstr = 7, dst = R.PA + col, ld = nX;
        else {
          const int code = R.rmap[col];
          a = code >> 5, src = R.Bb + (a * 7) * NP + ((code & 31) - 7), sstr = NP, dst = R.PB + (col - idt), ld = nu;
        }
        double cf[7];
#pragma unroll
        for (int m = 0; m < 7; ++m) cf[m] = src[m * sstr];
        for (int r = rg; r < nX; r += nrg) {
          const double* Pr = R.P + r * nX + 7 * a;
          double s = 0;
#pragma unroll
          for (int m = 0; m < 7; ++m) s += Pr[m] * cf[m];
          dst[r * ld] = s;
        }
      }
    }
    for (int it = ctx.nt - 1 - ctx.tid; it < 2 * nX; it += ctx.nt) {
      const int r = it >> 1;
      const double* Pr = R.P + r * nX;
      const double* vec = (it & 1) ? R.cb : R.db;
      double s0 = (it & 1) ? R.p[r] : Pr[idt], s1 = 0;
      int m = 0;
      for (; m + 1 < idt; m += 2) s0 += Pr[m] * vec[m], s1 += Pr[m + 1] * vec[m + 1];
      if (m < idt) s0 += Pr[m] * vec[m];
      if (it & 1) R.pc[r] = s0 + s1;
      else R.PA[r * nX + idt] = s0 + s1;
    }
    cta_sync(ctx);
    const int nc = nX + 1;
    const bool fast = ldl_fast_shape(ctx, nu, nc);
#if OBCA_RIC_PREFETCH
    if (i > 0) ric_input_fetch<SM>(ctx, L, W, R, i - 1);  // the assembly of this stage is done with the buffers: next stage's inputs
#endif
    if (fast) {
      // products and L D L' factorisation in registers, one barrier per pivot; F <- L, Gm <- Khat, K <- -D^-1 Khat
      stage_ldl_fused(ctx, R, nu, nX, ok, (LdlBuf*)ctx.ldl);
      cta_sync(ctx);
      prof_mark(ctx, 14);
    } else {
      // F = R + B'PB ; Gm = [S + B'PA | r + B'pc]
      for (int it = ctx.tid; it < nu * (nu + nX + 1); it += ctx.nt) {
        int u = it / (nu + nX + 1), col = it % (nu + nX + 1), a = 0;
        while (u >= R.uoff[a + 1]) ++a;
        int j = u - R.uoff[a];
        const double* Bj = R.Bb + (a * 7) * NP + j;
        double s = 0;
        if (col < nu) {
          for (int m = 0; m < 7; ++m) s += Bj[m * NP] * R.PB[(7 * a + m) * nu + col];
          R.F[u * nu + col] = R.R[u * nu + col] + s;
        } else if (col < nu + nX) {
          int c = col - nu;
          for (int m = 0; m < 7; ++m) s += Bj[m * NP] * R.PA[(7 * a + m) * nX + c];
          R.Gm[u * (nX + 1) + c] = R.S[u * nX + c] + s;
        } else {
          for (int m = 0; m < 7; ++m) s += Bj[m * NP] * R.pc[7 * a + m];
          R.Gm[u * (nX + 1) + nX] = R.r[u] + s;
        }
      }
      cta_sync(ctx);
      prof_mark(ctx, 13);
#if defined(__CUDA_ARCH__)
      riccati_factor_generic(ctx, R, nu, nc, i, ok);
#else
      stage_ldl_scalar(R.F, nu, R.Gm, nc, R.K, R.invd, R.R, ok);  // the emulation runs the L D L' form (one thread)
#endif
    }
    const bool ldl_form = fast || OBCA_EMU_LDL;
    // stage record for the forward pass: Ks transposed [nc][nUmax] (lanes read consecutive controls) and the unit L row major
    // [nUmax][nUmax]; the forward pass computes u = L^-T (Ks [x; 1]).  The generic path stores the gains as Ks and L = I.
    {
      double* Kg = W.RK + (size_t)i * kstride;
      double* Lg = Kg + (size_t)nc * nUmax;
      if (ctx.nt >= 64 && (ctx.nt & (ctx.nt - 1)) == 0) {
        // lanes over consecutive controls (coalesced global stores), thread groups over the columns / rows: no run-time divisions
        const int lu = nu <= 32 ? 5 : 6, u = ctx.tid & ((1 << lu) - 1), g0 = ctx.tid >> lu, ng = ctx.nt >> lu;
        if (u < nu) {
          for (int col = g0; col < nc; col += ng) Kg[col * nUmax + u] = R.K[u * nc + col];
          for (int r = g0; r < nu; r += ng)
            if (u < r) Lg[r * nUmax + u] = ldl_form ? R.F[r * nu + u] : 0.0;
        }
      } else {
        for (int it = ctx.tid; it < nu * nc; it += ctx.nt) {
          const int u = it / nc, col = it % nc;
          Kg[col * nUmax + u] = R.K[it];
        }
        for (int it = ctx.tid; it < nu * nu; it += ctx.nt) {
          const int r = it / nu, m = it % nu;
          if (m < r) Lg[r * nUmax + m] = ldl_form ? R.F[it] : 0.0;
        }
      }
    }
    // P <- Q + A'PA + Gm'K ; p <- q + A'pc + Gm'k, written straight into the cost-to-go of the next stage (shared memory) and into
    // the global copy the multiplier recovery reads: nothing of this stage reads the old P any more (PA, PB, pc hold its products).
    // P is symmetric: only col >= r is computed (rows r and nX-1-r share one strip of nX+1 items) and mirrored; item nX of
    // the strips < nX is the gradient entry.  The dense dt row only occurs in the items (dt, dt) and p[dt].
    {
      double* P0 = W.RP + (size_t)i * pstride;
      const int np1 = nX + 1, nstrip = (nX + 1) / 2;
      auto item = [&](int r, int col) {
        double s = 0;
        const double* rhs = col < nX ? R.PA + col : R.pc;  // column of PA (stride nX) or pc (stride 1)
        const int rs = col < nX ? nX : 1;
        if (r < idt) {
          const int a = r / 7, rr = r % 7;
          for (int m = 0; m < 7; ++m) s += R.Ab[a * 49 + m * 7 + rr] * rhs[(7 * a + m) * rs];
        } else {
          s = rhs[idt * rs];
          for (int m = 0; m < idt; ++m) s += R.db[m] * rhs[m * rs];
        }
        double s1 = 0;
        int m = 0;
        for (; m + 1 < nu; m += 2) s += R.Gm[m * np1 + r] * R.K[m * np1 + col], s1 += R.Gm[(m + 1) * np1 + r] * R.K[(m + 1) * np1 + col];
        if (m < nu) s += R.Gm[m * np1 + r] * R.K[m * np1 + col];
        s += s1;
        if (col < nX) {
          const double v = R.Q[r * nX + col] + s;
          R.P[r * nX + col] = v, R.P[col * nX + r] = v;
          P0[r * nX + col] = v, P0[col * nX + r] = v;
        } else {
          const double v = R.q[r] + s;
          R.p[r] = v, P0[nX * nX + r] = v;
        }
      };
      // strip p holds rows p and nX-1-p: np1 items.  Threads form a (strip group) x (item) grid, 32 or 64 items wide, so that no index
      // needs a division by a run-time value; the nX gradient entries of the strips' complement go to the last threads
      const bool grid2d = ctx.nt >= 64 && (ctx.nt & (ctx.nt - 1)) == 0;
      const int lce = np1 <= 32 ? 5 : 6;
      const int e0 = grid2d ? (ctx.tid & ((1 << lce) - 1)) : 0, pg = grid2d ? ctx.tid >> lce : 0, npg = grid2d ? ctx.nt >> lce : 1;
      for (int p = pg; p < nstrip; p += npg)
        for (int e = e0; e < np1; e += grid2d ? np1 : 1) {
          int r, col;
          if (e < nX - p) r = p, col = p + e;
          else {
            r = nX - 1 - p, col = r + (e - (nX - p));
            if (r == p) continue;  // middle row of an odd dimension: already covered
          }
          item(r, col);
        }
      for (int r = ctx.nt - 1 - ctx.tid; r < nX; r += ctx.nt) item(r, nX);
    }
    cta_sync(ctx);
  }
}

// Forward pass.  The stage record (Ks', L) and the dynamics rows of the T maps of stage i + 1 are fetched into shared memory
// with asynchronous copies (cp.async, two buffers) while stage i is computed; the state stays in shared memory (ping-pong).
// Per stage: warp 0 computes the controls  u = L^-T (Ks [x; 1])  -- one lane per control, its column of L in registers, the
// back-substitution runs on shuffles --, then all threads advance the state through the block dynamics.  Two CTA barriers per
// stage and no exposed global-memory latency (round 1: gains read from global memory inside the loop, 7 k cycles per stage).
OBCA_HD size_t ric_fwd_buffer_doubles(const Lay& L) { return ric_record_doubles(L) + (size_t)L.V * 7 * (NRED + 1); }

template <bool SM>
OBCA_HD void ric_fwd_fetch(const Ctx& ctx, const Lay& L, const Scratch& W, int i, double* buf) {
  const size_t kstride = ric_record_doubles(L);
  const double* Kg = W.RK + (size_t)i * kstride;
  double* fT = buf + kstride;
#if defined(__CUDA_ARCH__)
  if (SM) {
    for (int it = ctx.tid; it < (int)kstride; it += ctx.nt)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(buf + it)), "l"(Kg + it) : "memory");
  } else
#endif
  {
    for (int it = ctx.tid; it < (int)kstride; it += ctx.nt) buf[it] = Kg[it];
  }
  for (int it = ctx.tid; it < L.V * 7 * (NRED + 1); it += ctx.nt) {
    const int a = it / (7 * (NRED + 1)), r = (it / (NRED + 1)) % 7, cc = it % (NRED + 1);
    const bool active = i < L.N[a];
    const double* T = W.TT + (size_t)(a * L.Nmax + (active ? i : 0)) * (NW * NRED + NW);
    const double* src = cc < NRED ? T + (28 + r) * NRED + cc : T + NW * NRED + 28 + r;
#if defined(__CUDA_ARCH__)
    if (SM) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(fT + it)), "l"(src), "r"(active ? 8 : 0) : "memory");
      continue;
    }
#endif
    fT[it] = active ? *src : 0.0;
  }
}

template <bool SM>
OBCA_HDN void riccati_forward(const Ctx& ctx, const Lay& L, const Scratch& W, double* RW, int* ok) {
  assume_scratch(W);
  const int nX = L.nX, nUmax = L.nU, idt = 7 * L.V, V = L.V, nc = nX + 1;
  const size_t pstride = (size_t)nX * nX + nX, kstride = ric_record_doubles(L), bstride = ric_fwd_buffer_doubles(L);
  RicWork R;
  ric_carve<SM>(R, L, RW);  // R.npt (free directions of every block) is still valid; the stage matrices P .. K are dead and serve as scratch
  double* fb0 = R.P;
  double* Xs = fb0 + 2 * bstride;   // [2][nX] state ping-pong
  double* fU = Xs + 2 * nX;         // [nUmax] compact controls of the stage
  double* X = W.RX;
  double* U = W.RX + (size_t)(L.Nmax + 1) * nX;
  ric_fwd_fetch<SM>(ctx, L, W, 0, fb0);
  if (ctx.tid == 0) {
    for (int a = 0; a < V; ++a)
      for (int q = 0; q < NZ; ++q) Xs[7 * a + q] = -W.c[L.YINIT(a, q)];
    const double* P0 = W.RP;
    const double* p0 = P0 + nX * nX;
    double s = p0[idt];
    for (int m = 0; m < idt; ++m) s += P0[idt * nX + m] * Xs[m];
    double piv = P0[idt * nX + idt];
    if (!(piv > PIVOT_TOL * fmax(1.0, fabs(p0[idt])))) {
      *ok = 0;
      piv = 1.0;
    }
    Xs[idt] = -s / piv;
    for (int m = 0; m < nX; ++m) X[m] = Xs[m];
  }
  for (int i = 0; i < L.Nmax; ++i) {
    double* buf = fb0 + (size_t)(i & 1) * bstride;
    const double* fT = buf + kstride;
    const double* Xi = Xs + (size_t)(i & 1) * nX;
    double* Xn = Xs + (size_t)((i + 1) & 1) * nX;
    ric_input_wait();
    cta_sync(ctx);  // stage i's record has landed, Xi is complete, fU of the previous stage has been consumed
    if (i + 1 < L.Nmax) ric_fwd_fetch<SM>(ctx, L, W, i + 1, fb0 + (size_t)((i + 1) & 1) * bstride);
    int nu = 0;
    for (int a = 0; a < V; ++a) nu += R.npt[a * L.Nmax + i];
    const double* Lf = buf + (size_t)nc * nUmax;
#if defined(__CUDA_ARCH__)
    if (nu <= 32) {
      if ((ctx.tid >> 5) == 0) {
        const int lane = ctx.tid & 31;
        double Lc[32];  // this lane's column of L (rows below the diagonal)
#pragma unroll
        for (int r = 1; r < 32; ++r) Lc[r] = (r < nu && lane < r) ? Lf[r * nUmax + lane] : 0.0;
        double t = 0.0, t1 = 0.0;
        if (lane < nu) {
          t = buf[nX * nUmax + lane];
          int c = 0;
          for (; c + 1 < nX; c += 2) t += buf[c * nUmax + lane] * Xi[c], t1 += buf[(c + 1) * nUmax + lane] * Xi[c + 1];
          if (c < nX) t += buf[c * nUmax + lane] * Xi[c];
          t += t1;
        }
#pragma unroll
        for (int r = 31; r > 0; --r) {   // L' u = t: u_r is final once the rows above it have been applied
          if (r < nu) {                  // warp-uniform
            const double ur = __shfl_sync(0xffffffffu, t, r);
            t = fma(-Lc[r], ur, t);
          }
        }
        if (lane < nu) fU[lane] = t;
      }
    } else
#endif
    {
      // generic: t = Ks [x; 1] by all threads, then the back-substitution row by row
      for (int u = ctx.tid; u < nu; u += ctx.nt) {
        double t = buf[nX * nUmax + u];
        for (int c = 0; c < nX; ++c) t += buf[c * nUmax + u] * Xi[c];
        fU[u] = t;
      }
      cta_sync(ctx);
      for (int r = nu - 1; r > 0; --r) {
        const double ur = fU[r];
        for (int m = ctx.tid; m < r; m += ctx.nt) fU[m] -= Lf[r * nUmax + m] * ur;
        cta_sync(ctx);
      }
    }
    cta_sync(ctx);
    // compact index u -> (vehicle, slot), zero for the unused slots
    double* Ui = U + (size_t)i * nUmax;  // padded per-vehicle layout: NP slots per vehicle
    for (int t = ctx.tid; t < V * NP; t += ctx.nt) {
      const int a = t / NP, j = t % NP;
      int off = 0;
      for (int aa = 0; aa < a; ++aa) off += R.npt[aa * L.Nmax + i];
      Ui[t] = j < R.npt[a * L.Nmax + i] ? fU[off + j] : 0.0;
    }
    for (int r = ctx.tid; r < nX; r += ctx.nt) {
      double s;
      if (r == idt) s = Xi[idt];
      else {
        const int a = r / 7, rr = r % 7;
        s = 0;
        if (i < L.N[a]) {
          const double* Tr = fT + (size_t)(a * 7 + rr) * (NRED + 1);
          int off = 0;
          for (int aa = 0; aa < a; ++aa) off += R.npt[aa * L.Nmax + i];
          const int np = R.npt[a * L.Nmax + i];
          s = Tr[NRED] + Tr[IDT] * Xi[idt];
          for (int m = 0; m < 7; ++m) s += Tr[m] * Xi[7 * a + m];
          for (int j = 0; j < np; ++j) s += Tr[7 + j] * fU[off + j];
        }
      }
      Xn[r] = s;
      X[(size_t)(i + 1) * nX + r] = s;
    }
  }
  cta_sync(ctx);
}

// ------------------------------------------------------------------------------------------------
// [BACKSUB]
// ------------------------------------------------------------------------------------------------
OBCA_HDN void expand_primal(const Ctx& ctx, const Lay& L, const Scratch& W) {
  assume_scratch(W);
  const int nX = L.nX, nU = L.nU, idt = 7 * L.V;
  const double* X = W.RX;
  const double* U = W.RX + (size_t)(L.Nmax + 1) * nX;
  const double ddt = X[idt];
  for (int it = ctx.tid; it < L.V * L.Nmax; it += ctx.nt) {
    int a = it / L.Nmax, i = it % L.Nmax;
    if (i >= L.N[a]) continue;
    const double* T = W.TT + (size_t)(a * L.Nmax + i) * (NW * NRED + NW);
    const double* s0 = T + NW * NRED;
    double rc[NRED];
    for (int q = 0; q < 7; ++q) rc[q] = X[(size_t)i * nX + 7 * a + q];
    for (int q = 0; q < NP; ++q) rc[7 + q] = U[(size_t)i * nU + NP * a + q];
    rc[IDT] = ddt;
    for (int q = 0; q < NZ; ++q) W.dx[L.Z(a, q, i * NK)] = rc[q];
    for (int r = 0; r < NW; ++r) {
      double s = s0[r];
      for (int q = 0; q < NRED; ++q) s += T[r * NRED + q] * rc[q];
      W.dx[L.Z(a, r % NZ, i * NK + 1 + r / NZ)] = s;
    }
  }
  if (ctx.tid == 0) W.dx[L.oDT] = ddt;
}

// GN <- gn + Hn dz + sum_pairs Hc dpose_other + hd ddt   (stationarity residual before the J'dy terms)
OBCA_HDN void node_residual(const Ctx& ctx, const Lay& L, const Scratch& W) {
  assume_scratch(W);
  const double ddt = W.dx[L.oDT];
  for (int it = ctx.tid; it < L.V * L.Mv; it += ctx.nt) {
    int a = it / L.Mv, n = it % L.Mv;
    if (n >= L.M[a]) continue;
    const double* hn = W.HN + (size_t)a * 28 * L.Mv + n;
    double* gn = W.GN + (size_t)a * 7 * L.Mv + n;
    const double* hd = W.HD + (size_t)a * 7 * L.Mv + n;
    const size_t ns = L.Mv;
    double dz[NZ], out[NZ];
    for (int q = 0; q < NZ; ++q) dz[q] = W.dx[L.Z(a, q, n)];
    for (int r = 0; r < NZ; ++r) {
      double s = gn[r * ns] + hd[r * ns] * ddt;
      for (int m = 0; m < NZ; ++m) s += hn[sym(r, m) * ns] * dz[m];
      out[r] = s;
    }
    for (int p = 0; p < L.P; ++p) {
      if (n >= L.Mp[p]) continue;
      const double* ph = W.PH + (size_t)p * 27 * L.Mv + n;
      if (L.pa[p] == a) {
        int b = L.pb[p];
        for (int r = 0; r < 3; ++r)
          for (int m = 0; m < 3; ++m) out[r] += ph[sym(3 + m, r) * L.Mv] * W.dx[L.Z(b, m, n)];
      } else if (L.pb[p] == a) {
        int b = L.pa[p];
        for (int r = 0; r < 3; ++r)
          for (int m = 0; m < 3; ++m) out[r] += ph[sym(3 + r, m) * L.Mv] * W.dx[L.Z(b, m, n)];
      }
    }
    for (int q = 0; q < NZ; ++q) gn[q * ns] = out[q];
  }
}

// multiplier of constraint row r of block (a, i): collocation rows, terminal rows, then received implied rows
OBCA_HD double* block_row_multiplier(const Lay& L, const Scratch& W, int a, int i, int r) {
  if (r < 30) return &W.dy[L.YCOL(a, r % 5, i * NK + r / 5)];
  int nterm = (i == L.N[a] - 1) ? (4 + L.heading[a]) : 0;
  if (r < 30 + nterm) {
    int t = r - 30;
    return &W.dy[L.YTERM(a, L.heading[a] ? t : t + 1)];
  }
  return &W.QR[(size_t)(a * L.Nmax + i) * QRSZ + QR_NU + (r - 30 - nterm)];
}

// dy of the remaining row r (node-0, terminal or implied row) of block (a, i) += d, together with what it induces on the
// multipliers of the defining rows of the eliminated variables
OBCA_HD void add_row_multiplier(const Lay& L, const Scratch& W, int a, int i, int r, double d) {
  *block_row_multiplier(L, W, a, i, r) += d;
  const double* Q = W.QR + (size_t)(a * L.Nmax + i) * QRSZ;
  for (int k = 0; k < 5; ++k)  // y_D = z - W' y_R
    for (int g = 0; g < 3; ++g) *block_row_multiplier(L, W, a, i, (k + 1) * 5 + g) -= Q[(10 + g * 5 + k) * NC + r] * d;
  if (r >= 30) {
    *block_row_multiplier(L, W, a, i, 28) += Q[QR_BU + r] * d;
    *block_row_multiplier(L, W, a, i, 29) += Q[QR_BU + NC + r] * d;
  }
}

OBCA_HDN void recover_multipliers(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W) {
  assume_scratch(W);
  OBCA_ASSUME_STATIC(L, S);
  const int nX = L.nX;
  const size_t pstride = (size_t)nX * nX + nX;
  const double dt = W.x[L.oDT], idt = 1.0 / dt;
  // continuity multipliers = costate of the Riccati recursion
  for (int it = ctx.tid; it < L.V * L.Nmax; it += ctx.nt) {
    int a = it / L.Nmax, i = it % L.Nmax;
    if (i >= L.N[a] || i == 0) continue;
    const double* P = W.RP + (size_t)i * pstride;
    const double* pv = P + nX * nX;
    const double* Xi = W.RX + (size_t)i * nX;
    for (int q = 0; q < NZ; ++q) {
      double s = pv[7 * a + q];
      for (int m = 0; m < nX; ++m) s += P[(7 * a + q) * nX + m] * Xi[m];
      W.dy[L.YCONT(a, q, i)] = s;
    }
  }
  cta_sync(ctx);
  for (int it = ctx.tid; it < L.V * L.Nmax; it += ctx.nt) {
    int a = it / L.Nmax, i = it % L.Nmax;
    if (i >= L.N[a]) continue;
    int n0 = i * NK;
    const double* QRm = W.QR + (size_t)(a * L.Nmax + i) * QRSZ;
    const double* piv = QRm + QR_PIV;
    int rk = (int)QRm[QR_META], nr = (int)QRm[QR_META + 1];
    double v[NW], vt[NU2], dyr[NC];
    for (int r = 0; r < NW; ++r) v[r] = -W.GN[((size_t)a * 7 + r % NZ) * L.Mv + n0 + 1 + r / NZ];
    if (i < L.N[a] - 1)
      for (int q = 0; q < NZ; ++q) v[28 + q] -= W.dy[L.YCONT(a, q, i + 1)];
    // (1) control equations  -y_(k,c) + sum beta y = v_(a_k | w_k)  substituted into the (v, delta) equations
    for (int j = 1; j < NK; ++j)
      for (int q = 3; q < 5; ++q) {
        double sacc = v[(j - 1) * 7 + q];
        for (int k = 1; k < NK; ++k) sacc += S.cA[j][k] * idt * v[(k - 1) * 7 + q + 2];
        v[(j - 1) * 7 + q] = sacc;
      }
    // (2) z = B^-T v_P on the defining rows of (x, y, psi); reduced right-hand side v_U - N' z
    double zx[5], zy[5], zp[5], nd[5][5];
    for (int k = 0; k < 5; ++k) {
      const int n = n0 + 1 + k;
      const double psi = W.x[L.Z(a, 2, n)], vv = W.x[L.Z(a, 3, n)], tde = tan(W.x[L.Z(a, 4, n)]);
      nd[k][0] = cos(psi), nd[k][1] = sin(psi), nd[k][2] = tde / S.wb, nd[k][3] = vv * (1.0 + tde * tde) / S.wb, nd[k][4] = vv;
    }
    for (int k = 0; k < 5; ++k) {
      double sx = 0, sy = 0;
      for (int j = 0; j < 5; ++j) sx += S.cAi[j][k] * v[j * 7 + 0], sy += S.cAi[j][k] * v[j * 7 + 1];
      zx[k] = sx * dt, zy[k] = sy * dt;
    }
    for (int k = 0; k < 5; ++k) {
      double sp = 0;
      for (int j = 0; j < 5; ++j) sp += S.cAi[j][k] * (v[j * 7 + 2] - nd[j][4] * nd[j][1] * zx[j] + nd[j][4] * nd[j][0] * zy[j]);
      zp[k] = sp * dt;
    }
    for (int k = 0; k < 5; ++k) {
      vt[2 * k] = v[k * 7 + 3] + nd[k][0] * zx[k] + nd[k][1] * zy[k] + nd[k][2] * zp[k];
      vt[2 * k + 1] = v[k * 7 + 4] + nd[k][3] * zp[k];
    }
    apply_q(QRm, rk, vt, true);
    // R dy = vt[0:rk] on the staircase; dropped (dependent) rows keep dy = 0 here
    for (int r = 0; r < nr; ++r) dyr[r] = 0;
    for (int ii = rk - 1; ii >= 0; --ii) {
      int j = (int)piv[ii];
      double s = vt[ii];
      for (int m = ii + 1; m < rk; ++m) {
        int jm = (int)piv[m];
        s -= QRm[ii * NC + jm] * dyr[jm];
      }
      dyr[j] = s / QRm[ii * NC + j];
    }
    // defining rows of (x, y, psi): y_D = z - W' y_R
    for (int k = 0; k < 5; ++k) {
      double sx = zx[k], sy = zy[k], sp = zp[k];
      for (int r = 0; r < nr; ++r) {
        if (r >= 5 && r < 30) continue;
        sx -= QRm[(10 + k) * NC + r] * dyr[r], sy -= QRm[(15 + k) * NC + r] * dyr[r], sp -= QRm[(20 + k) * NC + r] * dyr[r];
      }
      dyr[(k + 1) * 5 + 0] = sx, dyr[(k + 1) * 5 + 1] = sy, dyr[(k + 1) * 5 + 2] = sp;
    }
    // defining rows of the controls
    for (int k = 1; k < NK; ++k)
      for (int c = 5; c < 7; ++c) {
        double s = -v[(k - 1) * 7 + c];
        if (k == NK - 1)
          for (int r = 30; r < nr; ++r) s += QRm[QR_BU + (c - 5) * NC + r] * dyr[r];
        dyr[k * 5 + c - 2] = s;
      }
    for (int r = 0; r < nr; ++r) *block_row_multiplier(L, W, a, i, r) = dyr[r];
    if (i == L.N[a] - 1 && !L.heading[a]) W.dy[L.YTERM(a, 0)] = 0.0;
    if (i == 0) {
      // node 0 of the first interval: gn0 + G0' dycol + dyinit = 0
      double psi = W.x[L.Z(a, 2, 0)], vv = W.x[L.Z(a, 3, 0)], de = W.x[L.Z(a, 4, 0)];
      double cs = cos(psi), sn = sin(psi), tde = tan(de), sec2 = 1.0 + tde * tde;
      double gy[NZ] = {0, 0, 0, 0, 0, 0, 0};
      for (int q = 0; q < 5; ++q)
        for (int k = 0; k < NK; ++k) gy[q] += S.cA[0][k] * idt * dyr[k * 5 + q];
      gy[2] -= dyr[0] * (-vv * sn) + dyr[1] * (vv * cs);
      gy[3] -= dyr[0] * cs + dyr[1] * sn + dyr[2] * tde / S.wb;
      gy[4] -= dyr[2] * vv * sec2 / S.wb;
      gy[5] -= dyr[3];
      gy[6] -= dyr[4];
      for (int q = 0; q < NZ; ++q) W.dy[L.YINIT(a, q)] = -W.GN[((size_t)a * 7 + q) * L.Mv] - gy[q];
    }
  }
  cta_sync(ctx);
  // implied rows: their multipliers nu (known at block i once block i is final) flow to the dependent rows of
  // block i+1 and to the continuity multiplier between the two blocks.  One thread per vehicle, upwards.
  for (int a = ctx.tid; a < L.V; a += ctx.nt) {
    for (int i = 0; i + 1 < L.N[a]; ++i) {
      const double* nu = W.QR + (size_t)(a * L.Nmax + i) * QRSZ + QR_NU;
      const double* Qn = W.QR + (size_t)(a * L.Nmax + i + 1) * QRSZ;
      int ndrop = (int)Qn[QR_META + 2];
      for (int d = 0; d < ndrop; ++d) {
        const double* dr = Qn + QR_DROP + d * DRSZ;
        int j = (int)dr[0], slot = (int)dr[1], rkd = (int)dr[2];
        if (slot < 0) continue;
        double nv = nu[slot];
        add_row_multiplier(L, W, a, i + 1, j, nv);
        for (int q = 0; q < rkd; ++q) add_row_multiplier(L, W, a, i + 1, (int)Qn[QR_PIV + q], -dr[3 + q] * nv);
        for (int q = 0; q < NZ; ++q) W.dy[L.YCONT(a, q, i + 1)] += nv * dr[3 + 35 + q];
      }
    }
  }
}

OBCA_HDN void local_backsub(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W) {
  assume_scratch(W);
  OBCA_ASSUME_STATIC(L, S);
  for (int it = ctx.tid; it < L.V * L.Mv; it += ctx.nt) {
    int a = it / L.Mv, n = it % L.Mv;
    if (n >= L.M[a]) continue;
    double dp[3] = {W.dx[L.Z(a, 0, n)], W.dx[L.Z(a, 1, n)], W.dx[L.Z(a, 2, n)]};
    for (int j = 0; j < L.O; ++j) obs_block_backsub(L, W, a, n, j, dp);
    int q = tube_set_at(L, a, n);
    if (q >= 1) {
      double psi = W.x[L.Z(a, 2, n)], cs = cos(psi), sn = sin(psi);
      for (int r = 0; r < 8; ++r) {
        const double* t = S.tube_row(L, a, q, r / 4, r % 4);
        double gr[3] = {t[0], t[1], r < 4 ? 0.0 : S.wb * (-t[0] * sn + t[1] * cs)};
        double dts = W.c[L.YTUBE(a, q - 1, r)] - (gr[0] * dp[0] + gr[1] * dp[1] + gr[2] * dp[2]);
        W.dx[L.TS(a, q - 1, r)] = dts;
        W.dy[L.YTUBE(a, q - 1, r)] = W.sig[L.TS(a, q - 1, r)] * dts + W.gphi[L.TS(a, q - 1, r)];
      }
    }
  }
  for (int it = ctx.tid; it < L.nPairNodes; it += ctx.nt) {  // compact index over the existing (pair, node) blocks
    int p = 0, n = it;
    while (n >= L.Mp[p]) n -= L.Mp[p], ++p;
    int a = L.pa[p], b = L.pb[p];
    double dp[6] = {W.dx[L.Z(a, 0, n)], W.dx[L.Z(a, 1, n)], W.dx[L.Z(a, 2, n)], W.dx[L.Z(b, 0, n)], W.dx[L.Z(b, 1, n)], W.dx[L.Z(b, 2, n)]};
    pair_block_backsub(L, W, p, n, dp);
  }
}

// Newton step at (W.x, W.y) for the right-hand side (W.gphi, W.c) and diagonal W.sig (Sigma + delta_w).
// Returns ok = 1 when the inertia is correct.  All threads of the CTA must call it.
OBCA_HDN int kkt_solve(const Ctx& ctx, const Lay& L, const Stat& S, const Scratch& W, double* RW, int* ok_shared) {
  if (ctx.tid == 0) *ok_shared = 1;
  cta_sync(ctx);
  prof_mark(ctx, 11);
  pair_eliminate(ctx, L, S, W, ok_shared);
  cta_sync(ctx);
  prof_mark(ctx, 2);
  double hdtdt;
  node_assemble(ctx, L, S, W, ok_shared, &hdtdt);
  cta_sync(ctx);
  prof_mark(ctx, 3);
  interval_nullspace(ctx, L, S, W, ok_shared, ok_shared + 1, RW);
  cta_sync(ctx);
  prof_mark(ctx, 4);
  interval_cross(ctx, L, W, RW);
  cta_sync(ctx);
  prof_mark(ctx, 5);
  // the Riccati arena: shared memory, or the per-slot global arena when the stage matrices of V > 4 vehicles do not fit
  if (W.ricg) riccati_backward<false>(ctx, L, W, W.ricg, hdtdt, ok_shared);
  else riccati_backward<true>(ctx, L, W, RW, hdtdt, ok_shared);
  cta_sync(ctx);
  prof_mark(ctx, 6);
  int ok = *ok_shared;
  if (!ok) return 0;
  if (W.ricg) riccati_forward<false>(ctx, L, W, W.ricg, ok_shared);
  else riccati_forward<true>(ctx, L, W, RW, ok_shared);
  cta_sync(ctx);
  prof_mark(ctx, 7);
  ok = *ok_shared;
  if (!ok) return 0;
  expand_primal(ctx, L, W);
  cta_sync(ctx);
  node_residual(ctx, L, W);
  cta_sync(ctx);
  prof_mark(ctx, 8);
  recover_multipliers(ctx, L, S, W);
  cta_sync(ctx);
  prof_mark(ctx, 9);
  local_backsub(ctx, L, S, W);
  cta_sync(ctx);
  prof_mark(ctx, 10);
  return 1;
}

}  // namespace obca
