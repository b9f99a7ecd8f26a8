// obca_ws.h -- dual warm starts in closed form.
//
// The reference obtains the initial obstacle duals from Vehicle.dual_ws (confrez/control/vehicle.py:233-296) and the
// initial pair duals from MultiVehiclePlanner.joint_dual_ws (confrez/control/multi_vehicle_planner.py:208-341): one
// IPOPT call each that maximises, per (node, obstacle) or (pair, node), the dual of the distance between two convex
// polytopes with the poses fixed.  For the 4-face polytopes of the reference scenarios that problem has a closed form:
// with w the unit vector between the closest points (the least-penetration axis when the shapes overlap)
//
//     obstacle:  lambda = max(0, A w),          mu = max(0, -G R' w)         (w from the obstacle towards the body)
//     pair:      s = w (from b towards a),      lambda_ab = max(0, G R_a'(-s)),  lambda_ba = max(0, G R_b' s)
//
// One thread per (instance, vehicle, node, obstacle) / (instance, pair, node); same candidate order as the host
// statement in control/warmstart.py so that ties between equally close features resolve identically.
#pragma once

namespace obca {

// vertices of {x: A x <= b}, 4 faces listed around the boundary: vertex i = face i  /\  face i+1
OBCA_HD void ws_rect_vertices(const double A[4][2], const double b[4], double V[4][2]) {
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const double det = A[i][0] * A[j][1] - A[i][1] * A[j][0];
    V[i][0] = (b[i] * A[j][1] - A[i][1] * b[j]) / det;
    V[i][1] = (A[i][0] * b[j] - b[i] * A[j][0]) / det;
  }
}

// closest point on segment ab to p
OBCA_HD void ws_point_segment(const double p[2], const double a[2], const double b[2], double out[2]) {
  const double ab0 = b[0] - a[0], ab1 = b[1] - a[1];
  double t = ((p[0] - a[0]) * ab0 + (p[1] - a[1]) * ab1) / fmax(ab0 * ab0 + ab1 * ab1, 1e-300);
  t = fmin(fmax(t, 0.0), 1.0);
  out[0] = a[0] + t * ab0, out[1] = a[1] + t * ab1;
}

// unit vector pointing from the convex quad Q towards the convex quad P: direction between the closest points, or the
// least-penetration face normal when they overlap.  NP, NQ: outward unit face normals.
OBCA_HD void ws_direction(const double P[4][2], const double NPn[4][2], const double Q[4][2], const double NQn[4][2], double w[2]) {
  double best = INFINITY, bp[2] = {0, 0}, bq[2] = {0, 0};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double c[2];
      ws_point_segment(P[i], Q[j], Q[(j + 1) & 3], c);  // vertex i of P against edge j of Q
      double d = sqrt((P[i][0] - c[0]) * (P[i][0] - c[0]) + (P[i][1] - c[1]) * (P[i][1] - c[1]));
      if (d < best) best = d, bp[0] = P[i][0], bp[1] = P[i][1], bq[0] = c[0], bq[1] = c[1];
      ws_point_segment(Q[i], P[j], P[(j + 1) & 3], c);  // vertex i of Q against edge j of P
      d = sqrt((Q[i][0] - c[0]) * (Q[i][0] - c[0]) + (Q[i][1] - c[1]) * (Q[i][1] - c[1]));
      if (d < best) best = d, bp[0] = c[0], bp[1] = c[1], bq[0] = Q[i][0], bq[1] = Q[i][1];
    }
  // separating-axis test: separation along the face normals of Q and of P (positive = disjoint)
  double sepQ = -INFINITY, sepP = -INFINITY;
  int iq = 0, ip = 0;
  for (int f = 0; f < 4; ++f) {
    double mnP = INFINITY, mxQ = -INFINITY, mnQ = INFINITY, mxP = -INFINITY;
    for (int v = 0; v < 4; ++v) {
      mnP = fmin(mnP, NQn[f][0] * P[v][0] + NQn[f][1] * P[v][1]);
      mxQ = fmax(mxQ, NQn[f][0] * Q[v][0] + NQn[f][1] * Q[v][1]);
      mnQ = fmin(mnQ, NPn[f][0] * Q[v][0] + NPn[f][1] * Q[v][1]);
      mxP = fmax(mxP, NPn[f][0] * P[v][0] + NPn[f][1] * P[v][1]);
    }
    if (mnP - mxQ > sepQ) sepQ = mnP - mxQ, iq = f;
    if (mnQ - mxP > sepP) sepP = mnQ - mxP, ip = f;
  }
  const double sep = fmax(sepQ, sepP);
  const double d0 = bp[0] - bq[0], d1 = bp[1] - bq[1];
  const double n = sqrt(d0 * d0 + d1 * d1);
  if (sep <= 1e-9 || n <= 1e-9) {
    if (sepQ >= sepP) w[0] = NQn[iq][0], w[1] = NQn[iq][1];
    else w[0] = -NPn[ip][0], w[1] = -NPn[ip][1];
  } else
    w[0] = d0 / fmax(n, 1e-300), w[1] = d1 / fmax(n, 1e-300);
}

// body polytope {G q <= g} placed at pose (x, y, psi): vertices and outward normals in the world frame
OBCA_HD void ws_body(const double G[4][2], const double g[4], double x, double y, double psi, double V[4][2], double N[4][2]) {
  double Vb[4][2];
  ws_rect_vertices(G, g, Vb);
  const double c = cos(psi), s = sin(psi);
  for (int k = 0; k < 4; ++k) {
    V[k][0] = c * Vb[k][0] - s * Vb[k][1] + x;
    V[k][1] = s * Vb[k][0] + c * Vb[k][1] + y;
    N[k][0] = c * G[k][0] - s * G[k][1];
    N[k][1] = s * G[k][0] + c * G[k][1];
  }
}

// one (node, obstacle): lam[4], mu[4]
OBCA_HD void ws_obstacle_duals(const Stat& S, int j, double x, double y, double psi, double* lam, double* mu) {
  double Pb[4][2], Nb[4][2], Qo[4][2], w[2];
  ws_body(S.G, S.g, x, y, psi, Pb, Nb);
  ws_rect_vertices(S.obsA[j], S.obsb[j], Qo);
  ws_direction(Pb, Nb, Qo, S.obsA[j], w);
  const double c = cos(psi), s = sin(psi);
  const double wb0 = c * w[0] + s * w[1], wb1 = -s * w[0] + c * w[1];  // R' w
  for (int r = 0; r < 4; ++r) {
    lam[r] = fmax(0.0, S.obsA[j][r][0] * w[0] + S.obsA[j][r][1] * w[1]);
    mu[r] = fmax(0.0, -(S.G[r][0] * wb0 + S.G[r][1] * wb1));
  }
}

// one (pair, node): lam_ab[4], lam_ba[4], s[2]
OBCA_HD void ws_pair_duals(const Stat& S, double xa, double ya, double pa, double xb, double yb, double pb, double* lam, double* mu, double* sv) {
  double Pa[4][2], Na[4][2], Pb[4][2], Nb[4][2], s[2];
  ws_body(S.G, S.g, xa, ya, pa, Pa, Na);
  ws_body(S.G, S.g, xb, yb, pb, Pb, Nb);
  ws_direction(Pa, Na, Pb, Nb, s);  // from b towards a
  const double ca = cos(pa), sa = sin(pa), cb = cos(pb), sb = sin(pb);
  const double a0 = -(ca * s[0] + sa * s[1]), a1 = -(-sa * s[0] + ca * s[1]);  // R_a'(-s)
  const double b0 = cb * s[0] + sb * s[1], b1 = -sb * s[0] + cb * s[1];        // R_b' s
  for (int r = 0; r < 4; ++r) {
    lam[r] = fmax(0.0, S.G[r][0] * a0 + S.G[r][1] * a1);
    mu[r] = fmax(0.0, S.G[r][0] * b0 + S.G[r][1] * b1);
  }
  sv[0] = s[0], sv[1] = s[1];
}

// flat work items over node-major user arrays: z (B,V,Mmax,7) -> lam, mu (B,V,Mmax,O,4)
OBCA_HD void ws_obstacle_item(const Lay& L, const Stat& S, const double* z, double* lam, double* mu, size_t g) {
  const int O = L.O, Mmax = L.Mv;
  const int j = (int)(g % O);
  const size_t node = g / O;  // (b * V + a) * Mmax + n
  const int n = (int)(node % Mmax), a = (int)((node / Mmax) % L.V);
  double l[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
  if (n < L.M[a]) ws_obstacle_duals(S, j, z[node * 7 + 0], z[node * 7 + 1], z[node * 7 + 2], l, m);
  for (int r = 0; r < 4; ++r) lam[g * 4 + r] = l[r], mu[g * 4 + r] = m[r];
}

// z (B,V,Mmax,7) -> pair_lam, pair_mu (B,P,Mmax,4), pair_s (B,P,Mmax,2)
OBCA_HD void ws_pair_item(const Lay& L, const Stat& S, const double* z, double* pl, double* pm, double* ps, size_t g) {
  const int Mmax = L.Mv;
  const int n = (int)(g % Mmax), p = (int)((g / Mmax) % L.P);
  const size_t b = g / ((size_t)Mmax * L.P);
  double l[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, s[2] = {0, 0};
  if (n < L.Mp[p]) {
    const double* za = z + ((b * L.V + L.pa[p]) * Mmax + n) * 7;
    const double* zb = z + ((b * L.V + L.pb[p]) * Mmax + n) * 7;
    ws_pair_duals(S, za[0], za[1], za[2], zb[0], zb[1], zb[2], l, m, s);
  }
  for (int r = 0; r < 4; ++r) pl[g * 4 + r] = l[r], pm[g * 4 + r] = m[r];
  ps[g * 2] = s[0], ps[g * 2 + 1] = s[1];
}

}  // namespace obca
