"""Index maps between the oracle's private numbering (oracle/nlp.py) and the solver's internal flat layout
(obca_core.h ``Lay``), used by the parity tests to compare raw vectors (iterates, residuals, Newton steps)."""
import numpy as np


class DeviceLayout:
    def __init__(self, L):
        self.L = L

    def Z(self, a, c, n):
        return self.L["oZ"] + (a * 7 + c) * self.L["Mv"] + n

    def LAM(self, a, j, r, n):
        return self.L["oLAM"] + ((a * self.L["O"] + j) * 4 + r) * self.L["Mv"] + n

    def MU(self, a, j, r, n):
        return self.L["oMU"] + ((a * self.L["O"] + j) * 4 + r) * self.L["Mv"] + n

    def SD(self, a, j, n):
        return self.L["oSD"] + (a * self.L["O"] + j) * self.L["Mv"] + n

    def EL(self, a, j, n):
        return self.L["oEL"] + (a * self.L["O"] + j) * self.L["Mv"] + n

    def TS(self, a, q, r):
        return self.L["oTS"] + (a * (self.L["Smax"] - 1) + q) * 8 + r

    def PAIR(self, base, width, p, r, n):
        return self.L[base] + (p * width + r) * self.L["Mv"] + n

    def YCOL(self, a, c, n):
        return self.L["oYCOL"] + (a * 5 + c) * self.L["Mv"] + n

    def YCONT(self, a, c, i):
        return self.L["oYCONT"] + (a * 7 + c) * self.L["Nmax"] + i

    def YOBS(self, a, j, r, n):
        return self.L["oYOBS"] + ((a * self.L["O"] + j) * 4 + r) * self.L["Mv"] + n

    def YTUBE(self, a, q, r):
        return self.L["oYTUBE"] + (a * (self.L["Smax"] - 1) + q) * 8 + r

    def YPAIR(self, p, r, n):
        return self.L["oYPAIR"] + (p * 6 + r) * self.L["Mv"] + n


def build_maps(L, nlp):
    """Returns (ix, iy): ix[k] = device index of oracle variable k, iy[r] = device index of oracle row r."""
    D = DeviceLayout(L)
    V, O = nlp.prob.V, nlp.prob.O
    ix = np.full(nlp.n, -1, dtype=np.int64)
    iy = np.full(nlp.m, -1, dtype=np.int64)
    for a in range(V):
        M, N, S = nlp.M[a], nlp.N[a], int(nlp.prob.n_sets[a])
        n = np.arange(M)
        for c in range(7):
            ix[nlp.iz[a][:, c]] = D.Z(a, c, n)
        for j in range(O):
            for r in range(4):
                ix[nlp.ilam[a][:, j, r]] = D.LAM(a, j, r, n)
                ix[nlp.imu[a][:, j, r]] = D.MU(a, j, r, n)
                iy[nlp.r_obs[a][:, j, r]] = D.YOBS(a, j, r, n)
            ix[nlp.isd[a][:, j]] = D.SD(a, j, n)
            ix[nlp.iel[a][:, j]] = D.EL(a, j, n)
        for q in range(S - 1):
            for r in range(8):
                ix[nlp.its[a][q, r]] = D.TS(a, q, r)
                iy[nlp.r_tube[a][q, r]] = D.YTUBE(a, q, r)
        for c in range(7):
            iy[nlp.r_init[a][c]] = L["oYINIT"] + a * 7 + c
            iy[nlp.r_cont[a][:, c]] = D.YCONT(a, c, np.arange(1, N))
        for c in range(5):
            iy[nlp.r_col[a][:, c]] = D.YCOL(a, c, n)
        comps = ([0] if nlp.has_heading[a] else []) + [1, 2, 3, 4]
        for row, comp in zip(nlp.r_term[a], comps):
            iy[row] = L["oYTERM"] + a * 5 + comp
    for q in range(len(nlp.pairs)):
        m = nlp.Mp[q]
        n = np.arange(m)
        for r in range(4):
            ix[nlp.ipl[q][:, r]] = D.PAIR("oPL", 4, q, r, n)
            ix[nlp.ipm[q][:, r]] = D.PAIR("oPM", 4, q, r, n)
        for r in range(2):
            ix[nlp.ips[q][:, r]] = D.PAIR("oPS", 2, q, r, n)
        ix[nlp.ipsd[q]] = D.PAIR("oPSD", 1, q, 0, n)
        ix[nlp.ipsn[q]] = D.PAIR("oPSN", 1, q, 0, n)
        ix[nlp.ipel[q]] = D.PAIR("oPEL", 1, q, 0, n)
        for r in range(6):
            iy[nlp.r_pair[q][:, r]] = D.YPAIR(q, r, n)
    ix[nlp.idt] = L["oDT"]
    assert (ix >= 0).all() and (iy >= 0).all()
    assert len(np.unique(ix)) == nlp.n and len(np.unique(iy)) == nlp.m
    return ix, iy
