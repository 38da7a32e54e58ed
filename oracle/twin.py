"""Independent pure-Python twin of the routing path (second restatement, used to cross-check mr_oracle.c).

TEST INFRASTRUCTURE ONLY.  Written separately from mr_oracle.c, directly from the Fortran
(/root/reference/route/build/src: kwt_route.f90, irf_route.f90, basinUH.f90, accum_runoff.f90,
process_remap.f90:319-422, process_param.f90, gamma_func.f90, lake_route.f90:87-229,
network_topo.f90, process_ntopo.f90:176-187).  Plain Python floats are IEEE doubles evaluated
one operation at a time, so when both restatements follow the same operation order they agree to
the last bit; tests/test_oracle_twin.py asserts <= 1e-12 relative.  Slow: small networks only.
Fortran arrays declared (0:n) are Python lists indexed the same way; arrays declared (1:n) are
lists with a dummy element 0.
"""
from __future__ import annotations

import math
import struct
import sys

HUGE = sys.float_info.max
TINY = sys.float_info.min
EPS = sys.float_info.epsilon
MAXQPAR = 20


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


# ---------------------------------------------------------------- gamma_func.f90
def gammln(xx):
    coef = (76.18009172947146, -86.50532032941677, 24.01409824083091,
            -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5)
    x = xx
    tmp = x + 5.5
    tmp = (x + 0.5) * math.log(tmp) - tmp
    den = x + 1.0
    s = 0.0
    for c in coef:
        s += c / den
        den = den + 1.0
    return tmp + math.log(2.5066282746310005 * (1.000000000190015 + s) / x)


def gser(a, x):
    if x == 0.0:
        return 0.0
    ap, summ = a, 1.0 / a
    dl = summ
    for _ in range(100):
        ap += 1.0
        dl = dl * x / ap
        summ += dl
        if abs(dl) < abs(summ) * EPS:
            break
    return summ * math.exp(-x + a * math.log(x) - gammln(a))


def gcf(a, x):
    if x == 0.0:
        return 1.0
    fpmin = TINY / EPS
    b = x + 1.0 - a
    c = 1.0 / fpmin
    d = 1.0 / b
    h = d
    for i in range(1, 101):
        an = -i * (i - a)
        b += 2.0
        d = an * d + b
        if abs(d) < fpmin:
            d = fpmin
        c = b + an / c
        if abs(c) < fpmin:
            c = fpmin
        d = 1.0 / d
        dl = d * c
        h = h * dl
        if abs(dl - 1.0) <= EPS:
            break
    return math.exp(-x + a * math.log(x) - gammln(a)) * h


def gammp(a, x):
    return gser(a, x) if x < a + 1.0 else 1.0 - gcf(a, x)


# ---------------------------------------------------------------- process_param.f90
def basin_uh(dt, fshape, tscale):
    cum = gammp(fshape, dt / tscale)
    if cum > 0.999:
        ntry = 1.999
    else:
        lo, hi = 1.0, 1000.0
        ntry = 0.5 * (lo + hi)
        for it in range(1, 101):
            cum = gammp(fshape, dt * ntry / tscale)
            if cum < 0.99:
                lo = ntry
            if cum > 0.999:
                hi = ntry
            if 0.99 < cum < 0.999:
                break
            ntry = 0.5 * (lo + hi)
            if it == 100:
                raise RuntimeError("cannot identify the maximum number of bins for the tdh")
    ntdh = int(math.ceil(ntry))
    ff, psave = [], 0.0
    for j in range(1, ntdh + 1):
        cum = gammp(fshape, (float(j) * dt) / tscale)
        ff.append(max(0.0, cum - psave))
        psave = cum
    s = 0.0
    for v in ff:
        s += v
    return [v / s for v in ff]


def make_uh(length, dt, velo, diff):
    pi = 3.14159265359
    nT = 240
    thr1, thr2 = _f32(0.99999), _f32(0.9999)
    nsub = int(math.ceil(dt / 3600.0))
    fr = [0.0] * (nT + 1)
    for k in range(1, nsub + 1):
        fr[k] = 1.0 / nsub
    uhm = [0.0] * (nT + 1)
    inte, sec = 0.0, 0.0
    for ih in range(1, nT + 1):
        sec = sec + 3600.0
        if velo > 0.0:
            pot = ((velo * sec - length) ** 2) / (4.0 * diff * sec)
            hh = 0.0 if pot > 69.0 else 1.0 / (2.0 * math.sqrt(pi * diff * sec)) * length * math.exp(-pot)
        else:
            hh = 0.0
        uhm[ih] = hh
        inte = inte + hh
    if inte > 0.0:
        uhm = [v / inte for v in uhm]
    inte, last = 0.0, 1
    for ih in range(1, nT + 1):
        inte += uhm[ih]
        last = ih
        if inte > thr1:
            break
    inte, strt = 0.0, 1
    for ih in range(nT, 0, -1):
        inte += uhm[ih]
        strt = ih
        if inte > thr1:
            break
    uhq = [0.0] * (nT + 1)
    inte = 0.0
    for jh in range(1, nT + 1):
        q0 = 0.0
        for ih in range(strt, last + 1):
            if jh - ih > 0:
                if jh - ih <= nsub:
                    q0 = q0 + fr[jh - ih] * uhm[ih]
            else:
                break
        uhq[jh] = q0
        inte = inte + q0
    if inte > 0.0:
        uhq = [v / inte for v in uhq]
    inte = 0.0
    for ih in range(1, nT + 1):
        inte += uhq[ih]
        last = ih
        if inte > thr2:
            break
    uhq = [v / inte for v in uhq]
    ntdh = (last + nsub - 1) // nsub
    out = [0.0] * ntdh
    for jh in range(1, last + 1):
        out[(jh + nsub - 1) // nsub - 1] += uhq[jh]
    return out


# ---------------------------------------------------------------- KWT helpers
# -- hydraulic.f90 (trapezoid + floodplain above bankDepth); written independently of the C oracle ----------------
def hy_btop(y, b, zc, zf, bd):
    if y <= bd:
        return b + 2 * y * zc
    return (b + 2 * bd * zc) + zf * (y - bd) * 2


def hy_pwet(y, b, zc, zf, bd):
    if y <= bd:
        return b + 2 * y * math.sqrt(1 + zc * zc)
    return (b + 2 * bd * math.sqrt(1 + zc * zc)) + 2 * (y - bd) * math.sqrt(1 + zf * zf)


def hy_area(y, b, zc, zf, bd):
    if y <= bd:
        return y * (b + zc * y)
    return bd * (b + zc * bd) + (y - bd) * (hy_btop(y, b, zc, zf, bd) + hy_btop(bd, b, zc, zf, bd)) / 2.0


def hy_water_height(area, b, zc, zf, bd):
    a_bank = hy_area(bd, b, zc, zf, bd)
    if area > a_bank:
        bb = hy_btop(bd, b, zc, zf, bd)
        return bd + (-bb + math.sqrt(bb * bb - 4.0 * zf * (a_bank - area))) / (2.0 * zf)
    if zc == 0:
        return area / b
    return (-b + math.sqrt(b * b + 4.0 * area * zc)) / (2.0 * zc)


def hy_flow_depth(q, b, zc, s, n, zf, bd):
    """Normal depth by Newton-Raphson to 0.5 % (hydraulic.f90:299-420)."""
    if not q > 1.0e-50:
        return 0.0
    err, depth = 100.0, 0.0
    abf, pbf, bbf = hy_area(bd, b, zc, zf, bd), hy_pwet(bd, b, zc, zf, bd), hy_btop(bd, b, zc, zf, bd)
    qbf = abf * (abf / pbf) ** (2.0 / 3.0) * math.sqrt(s) / n
    if q < qbf:
        t = math.sqrt(s) / n / q
        c1 = t * t * t
        c2 = 2 * math.sqrt(zc * zc + 1.0)
        y0 = (1.0 / c1 / (b * b * b)) ** (1.0 / 5.0)
        while err > 0.005:
            a, bt, p = hy_area(y0, b, zc, zf, bd), hy_btop(y0, b, zc, zf, bd), hy_pwet(y0, b, zc, zf, bd)
            a2 = a * a
            a4 = a2 * a2
            a5 = a4 * a
            hh = c1 * a5 / (p * p) - 1.0
            dh = c1 * (5 * a4 * bt * p - 2 * c2 * a5) / (p * p * p)
            depth = y0 - hh / dh
            err = abs((depth - y0) / depth)
            y0 = depth
    else:
        y0 = bd + 2.0
        c1 = math.sqrt(s) / n / pbf ** (2.0 / 3.0)
        c2 = 2 * (zf / 2) ** (5.0 / 3.0) * math.sqrt(s) / n / (zf * zf + 1.0) ** (1.0 / 3.0)
        while err > 0.005:
            ye = y0 - bd
            hh = c1 * (abf + bbf * ye) ** (5.0 / 3.0) + c2 * ye ** (10.0 / 3.0) / ye ** (2.0 / 3.0) - q
            dh = c1 * (5.0 / 3.0) * bbf * (abf + bbf * ye) ** (2.0 / 3.0) + c2 * (10.0 / 3.0 - 2.0 / 3.0) * ye ** (5.0 / 3.0)
            depth = y0 - hh / dh
            err = abs((depth - y0) / depth)
            y0 = depth
    return depth


def _hy_sf(q, y, b, zc, n, zf, bd):
    a, p = hy_area(y, b, zc, zf, bd), hy_pwet(y, b, zc, zf, bd)
    t = q * n / a / (a / p) ** (2.0 / 3.0)
    return t * t


def hy_celerity(q, y, b, zc, s, n, zf, bd):
    if not y > 0.0:
        return 0.0
    return 5.0 / 3.0 * _hy_sf(q, y, b, zc, n, zf, bd) ** 0.3 * q ** 0.4 / hy_btop(y, b, zc, zf, bd) ** 0.4 / n ** 0.6


def hy_diffusivity(q, y, b, zc, s, n, zf, bd):
    if not y > 0.0:
        return 0.0
    return abs(q) / _hy_sf(q, y, b, zc, n, zf, bd) / hy_btop(y, b, zc, zf, bd) / 2.0


def solve_ade(length, prev, dt, q_up, ck, dk):
    """advection_diffusion.f90: implicit central-difference step with a Neumann outlet, Thomas algorithm.  Written on the
    matrix rows (sub / diag / super per row) rather than the reference's column-stored diagonals."""
    nm = len(prev)
    dx = length / (nm - 2)
    cd = dk * dt / (dx * dx)
    ca = ck * dt / dx
    sub = [0.0] * nm      # A[i][i-1]
    dia = [0.0] * nm
    sup = [0.0] * nm      # A[i][i+1]
    rhs = [0.0] * nm
    dia[0], rhs[0] = 1.0, q_up
    for i in range(1, nm - 1):
        sub[i] = -1.0 * ca - 2.0 * 1.0 * cd
        dia[i] = 2.0 + 4 * 1.0 * cd
        sup[i] = 1.0 * ca - 2.0 * 1.0 * cd
        rhs[i] = (0.0 * ca + 2.0 * 0.0 * cd) * prev[i - 1] + (2.0 - 4.0 * 0.0 * cd) * prev[i] - (0.0 * ca - 2.0 * 0.0 * cd) * prev[i + 1]
    sub[nm - 1], dia[nm - 1], rhs[nm - 1] = -1.0, 1.0, prev[nm - 1] - prev[nm - 2]
    for i in range(1, nm):
        c = sub[i] / dia[i - 1]
        dia[i] = dia[i] - c * sup[i - 1]
        rhs[i] = rhs[i] - c * rhs[i - 1]
    out = [0.0] * nm
    out[nm - 1] = rhs[nm - 1] / dia[nm - 1]
    for i in range(nm - 2, -1, -1):
        out[i] = (rhs[i] - sup[i] * out[i + 1]) / dia[i]
    return out


class Wave:
    __slots__ = ("QF", "TI", "TR", "RF")

    def __init__(self, QF, TI, TR, RF):
        self.QF, self.TI, self.TR, self.RF = QF, TI, TR, RF

    def copy(self):
        return Wave(self.QF, self.TI, self.TR, self.RF)


class RouteError(RuntimeError):
    def __init__(self, ierr, msg):
        super().__init__(f"ierr={ierr}: {msg}")
        self.ierr = ierr


def interp_rch(told, qold, t0, t1):
    """kwt_route.f90:1444-1622 for a single interval; told/qold are 0-based Python lists."""
    T = [None] + list(told)
    Q = [None] + list(qold)
    n = len(told)
    if T[1] > t0 or T[n] < t1:
        raise RouteError(1, "interp_rch/bad bounds")
    ibeg = 1
    for i in range(2, n + 1):
        if t0 <= T[i]:
            ibeg = i
            break
    iend = 1
    for i in range(1, n + 1):
        if t1 <= T[i]:
            iend = i
            break
    if t1 < T[ibeg]:
        sl = (Q[ibeg] - Q[ibeg - 1]) / (T[ibeg] - T[ibeg - 1])
        q0 = sl * (t0 - T[ibeg - 1]) + Q[ibeg - 1]
        q1 = sl * (t1 - T[ibeg - 1]) + Q[ibeg - 1]
        return 0.5 * (q0 + q1)
    ab = ae = am = 0.0
    if t0 < T[ibeg]:
        sl = (Q[ibeg] - Q[ibeg - 1]) / (T[ibeg] - T[ibeg - 1])
        q0 = sl * (t0 - T[ibeg - 1]) + Q[ibeg - 1]
        ab = (T[ibeg] - t0) * 0.5 * (q0 + Q[ibeg])
    if t1 < T[iend]:
        sl = (Q[iend] - Q[iend - 1]) / (T[iend] - T[iend - 1])
        q1 = sl * (t1 - T[iend - 1]) + Q[iend - 1]
        ae = (t1 - T[iend - 1]) * 0.5 * (Q[iend - 1] + q1)
    if ibeg < iend:
        for im in range(ibeg + 1, iend + 1):
            if im < iend or (im == iend and t1 == T[iend] and t0 < T[iend - 1]):
                am = am + (T[im] - T[im - 1]) * 0.5 * (Q[im - 1] + Q[im])
    return (ab + ae + am) / (t1 - t0)


def remove_rch(Q, T, Z):
    """kwt_route.f90:999-1123; returns thinned copies."""
    nprt = len(Q) - 1
    flg = [True] * (nprt + 1)
    err = [HUGE] * (nprt + 1)

    def itp(t0, q1, q2, t1, t2):
        return q1 + ((q2 - q1) / (t2 - t1)) * (t0 - t1)

    for i in range(1, nprt):
        err[i] = abs(itp(T[i], Q[i - 1], Q[i + 1], T[i - 1], T[i + 1]) - Q[i])
    while True:
        idx = [i for i in range(nprt + 1) if flg[i]]
        mprt = len(idx) - 1
        if mprt < MAXQPAR:
            break
        e = [err[i] for i in idx]
        isel = e.index(min(e))
        if idx[isel - 1] > 0:
            a, m, p = idx[isel - 2], idx[isel - 1], idx[isel + 1]
            err[m] = abs(itp(T[m], Q[a], Q[p], T[a], T[p]) - Q[m])
        if idx[isel + 1] < nprt:
            a, m, p = idx[isel - 1], idx[isel + 1], idx[isel + 2]
            err[m] = abs(itp(T[m], Q[a], Q[p], T[a], T[p]) - Q[m])
        flg[idx[isel]] = False
    return [Q[i] for i in idx], [T[i] for i in idx], [Z[i] for i in idx]


def kinwav_rch(K, xmx, t_start, t_end, q_in, t_in):
    """kwt_route.f90:1130-1439.  q_in/t_in: particles 1..NQ1 as 0-based lists.
    Returns (Q, TENTRY, T_EXIT, FROUTE) lists of length NQ2."""
    ni = len(q_in)
    if ni == 0:
        return [], [], [], []
    alfa = 5.0 / 3.0
    nn = ni
    MF = [None] + list(range(1, ni + 1))
    IX = [None] + list(range(1, ni + 1)) + [None]
    Q0 = [None] + list(q_in)
    Q1 = [None] + list(q_in) + [None]
    Q2 = [None] + list(q_in) + [None]
    T0 = [None] + list(t_in)
    T1 = [None] + list(t_in) + [None]
    WC = [None] + [alfa * K ** (1.0 / alfa) * q ** ((alfa - 1.0) / alfa) for q in q_in] + [None]
    if nn > 1:
        x = 0.0
        while True:
            xb = xmx
            ixb = 0
            for iw in range(2, nn + 1):
                jw = iw - 1
                if WC[iw] == 0.0 or WC[jw] == 0.0:
                    continue
                wd = 1.0 / WC[jw] - 1.0 / WC[iw]
                if wd == 0.0:
                    continue
                if WC[iw] == WC[jw]:
                    continue
                xxb = (T1[iw] - T1[jw]) / wd
                if xxb < x or xxb > xb:
                    continue
                xb = xxb
                ixb = iw
            if xb == xmx:
                break
            nn -= 1
            jxb = ixb - 1
            Q2[jxb] = max(Q2[jxb], Q2[ixb])
            Q1[jxb] = min(Q1[jxb], Q1[ixb])
            a2 = (Q2[jxb] / K) ** (1.0 / alfa)
            a1 = (Q1[jxb] / K) ** (1.0 / alfa)
            cm = (Q2[jxb] - Q1[jxb]) / (a2 - a1)
            T1[jxb] = T1[jxb] + xb / WC[jxb] - xb / cm
            WC[jxb] = cm
            for i in range(IX[ixb], ni + 1):
                MF[i] -= 1
            for i in range(ixb, nn + 1):
                IX[i], T1[i], WC[i], Q1[i], Q2[i] = IX[i + 1], T1[i + 1], WC[i + 1], Q1[i + 1], Q2[i + 1]
            x = xb
    oQ, oT, oX, oF = [], [], [], []

    def rupdate(qn, told, tnew):
        if len(oQ) + 1 > ni:
            raise RouteError(60, "RUPDATE/array bounds exceeded")
        if oX and tnew <= oX[-1]:
            tnew = oX[-1] + 1.0
        if not oX and tnew <= t_start:
            tnew = t_start + 1.0
        oQ.append(qn); oT.append(told); oX.append(tnew); oF.append(tnew < t_end)

    for ir in range(1, nn + 1):
        if WC[ir] < TINY:
            raise RouteError(20, "kinwav_rch/zero flow")
        texit = min(xmx / WC[ir] + T1[ir], HUGE)
        tnext = min(xmx / WC[ir + 1] + T1[ir + 1], HUGE) if ir < nn else HUGE
        if Q1[ir] != Q2[ir]:
            if texit < t_end:
                texit2 = min(texit + 1.0, texit + 0.5 * (min(tnext, t_end) - texit))
                if texit2 == texit:
                    raise RouteError(30, "TEXIT equals TEXIT2 in kinwav")
                rupdate(Q1[ir], T1[ir], texit)
                rupdate(Q2[ir], T1[ir], texit2)
            else:
                for jr in range(1, ni + 1):
                    if MF[jr] == ir:
                        rupdate(Q0[jr], T0[jr], texit)
        else:
            rupdate(Q1[ir], T1[ir], texit)
    return oQ, oT, oX, oF


# ---------------------------------------------------------------- the routing domain
class Twin:
    def __init__(self, net, params, opts):
        self.net, self.p, self.o = net, params, opts
        n = net.nRch
        self.n = n
        self.tc, self.lc = opts.conv()
        self.methods = [int(c) for c in opts.route_opt]      # 0 SUM, 1 IRF, 2 KWT, 3 KW, 4 MC, 5 DW
        seg = [int(v) for v in net.segId]
        id2ix = {}
        for i, s in enumerate(seg):
            id2ix.setdefault(s, i)
        self.downId = [int(v) for v in net.downSegId]
        self.down = [id2ix.get(d, -1) if d > 0 else -1 for d in self.downId]
        self.ups = [[] for _ in range(n)]
        for i in range(n):
            if self.down[i] >= 0:
                self.ups[self.down[i]].append(i)
        self.hrus = [[] for _ in range(n)]
        for hix, sid in enumerate(net.hruSegId):
            j = id2ix.get(int(sid), -1) if sid > 0 else -1
            if j >= 0:
                self.hrus[j].append(hix)
        # processing order
        indeg = [len(u) for u in self.ups]
        order = [i for i in range(n) if indeg[i] == 0]
        k = 0
        while k < len(order):
            d = self.down[order[k]]
            k += 1
            if d >= 0:
                indeg[d] -= 1
                if indeg[d] == 0:
                    order.append(d)
        assert len(order) == n
        self.order = order
        area = [float(a) for a in net.area]
        self.bas, self.tot, self.wgt = [0.0] * n, [0.0] * n, [None] * n
        for r in order:
            ups_a = 0.0
            for u in self.ups[r]:
                ups_a = ups_a + self.tot[u]
            b = 0.0
            for hh in self.hrus[r]:
                b += area[hh]
            self.bas[r] = b
            self.tot[r] = b + ups_a
            self.wgt[r] = [area[hh] / b for hh in self.hrus[r]]
        self.ngood = [len(self.ups[r]) if self.tot[r] > TINY else 0 for r in range(n)]
        self.length = [float(v) for v in net.length]
        self.slope = [max(float(v), 1.0e-6) for v in net.slope]
        self.width = [float(net.width[i]) if net.width is not None else params.wscale * math.sqrt(self.tot[i]) for i in range(n)]
        self.man_n = [float(net.man_n[i]) if net.man_n is not None else params.mann_n for i in range(n)]
        lk = opts.is_lake_sim and net.islake is not None
        self.islake = [bool(lk and net.islake[i] == 1) for i in range(n)]
        self.lakeinlet = [self.down[i] >= 0 and self.islake[self.down[i]] for i in range(n)]
        if opts.is_lake_sim:
            self.ltype = [1 if (not opts.lakeRegulate or net.lakeModelType is None) else int(net.lakeModelType[i]) for i in range(n)]
        self.ff = basin_uh(opts.dt, params.fshape, params.tscale)
        self.uh = None
        if 1 in self.methods:
            self.uh = []
            for i in range(n):
                u = make_uh(self.length[i], opts.dt, params.velo, params.diff)
                if self.islake[i]:
                    u = [1.0] + [0.0] * (len(u) - 1)
                self.uh.append(u)
            self.qf_irf = [[0.0] * len(u) for u in self.uh]
        # state
        self.QI = [0.0] * n
        self.QR0 = [0.0] * n
        self.QR1 = [0.0] * n
        self.qfut = [[0.0] * len(self.ff) for _ in range(n)]
        self.Q = {m: [0.0] * n for m in range(6)}
        self.V0 = {m: [0.0] * n for m in range(6)}
        self.V1 = {m: [0.0] * n for m in range(6)}
        self.INF = {m: [0.0] * n for m in range(6)}
        self.WB = {m: [0.0] * n for m in range(6)}
        # Euler schemes: channel geometry (process_ntopo.f90:174-203) and molecules (init_model_data.f90:386-393,463-497)
        fp = bool(getattr(opts, "floodplain", False))
        self.depth = [4.5000000682193786e-05 * math.sqrt(self.tot[i]) if fp else 100000.0 for i in range(n)]
        self.zc = [0.0] * n
        self.zf = [1000.0] * n
        self.storage = [hy_area(self.depth[i], self.width[i], self.zc[i], self.zf[i], self.depth[i]) * self.length[i] for i in range(n)]
        self.mol = {m: [[0.0] * k for _ in range(n)] for m, k in ((3, 20), (4, 2), (5, 20)) if m in self.methods}
        self.FLOOD = {m: [0.0] * n for m in (3, 4, 5)}
        self.ELE = {m: [0.0] * n for m in (3, 4, 5)}
        self.KW = [None] * n
        if 2 in self.methods and opts.is_lake_sim:
            for i in range(n):
                if self.islake[i]:
                    self.KW[i] = [Wave(-9999.0, -9999.0, -9999.0, False)]
        self.T0, self.T1 = 0.0, float(opts.dt)
        self.itime = 1
        self.has_ep = False
        self.wm_flux = self.wm_vol = None
        self.wm_jump = False
        self.WMA = {m: [0.0] * n for m in range(6)}      # REACH_WM_FLUX_actual per method
        self.qmod, self.blend, self.trend, self.obs = 0, 10, 1, None
        self.Qobs, self.Qelapsed = [0.0] * n, [0] * n             # init_model_data.f90:404-405
        self.Qerr = {m: [0.0] * n for m in range(6)}

    # -- data assimilation by direct insertion: main_route.f90:125-148, data_assimilation.f90:23-97 --------------
    def set_da(self, qmod_option=1, q_blend_period=10, q_err_trend=1):
        self.qmod, self.blend, self.trend = int(qmod_option), int(q_blend_period), int(q_err_trend)

    def set_obs(self, obs=None):
        self.obs = None if obs is None else [float(x) for x in obs]

    def _read_obs(self):
        if self.obs is not None:
            for j, q in enumerate(self.obs):
                if math.isnan(q) or q < 0:
                    continue
                self.Qobs[j] = q
                self.Qelapsed[j] = 0
        else:
            self.Qelapsed = [e + 1 for e in self.Qelapsed]
        self.obs = None

    def _finish(self, m, j, qup, qlat):
        """End of a river reach: direct insertion instead of the water balance when qmodOption = 1."""
        if self.qmod != 1:
            return self._wb(m, j, qup, qlat)
        el, blend = self.Qelapsed[j], self.blend
        if self.Qobs[j] > 0.0:
            self.Qerr[m][j] = self.Q[m][j] - self.Qobs[j]
        if el > blend:
            self.Qerr[m][j] = 0.0
        err = self.Qerr[m][j]
        corr = 0.0
        if el <= blend:
            if self.trend == 1:
                corr = err
            elif self.trend == 2:
                corr = err * (1.0 - float(el) / float(blend))
            elif self.trend == 3:
                x0, y0 = 0.25, _f32(0.90)
                k = math.log(1.0 / y0 - 1.0) / (blend / 2.0 - blend * x0)
                corr = err / (1.0 + math.exp(-k * (1.0 * el - blend / 2.0)))
            elif self.trend == 4:
                if err != 0.0:
                    k = math.log(0.1 / abs(err)) / (1.0 * blend)
                    corr = err * math.exp(k * el)
            else:
                raise RouteError(81, "direct_insertion/discharge error trend model must be 1(const),2(liear), or 3(logistic)")
        self.Q[m][j] = max(self.Q[m][j] - corr, 0.0)

    def set_wm(self, flux_wm=None, vol_wm=None, vol_jumpstart=False):
        self.wm_flux = None if flux_wm is None else [float(x) for x in flux_wm]
        self.wm_vol = None if (vol_wm is None or not self.o.is_lake_sim) else [float(x) for x in vol_wm]
        self.wm_jump = bool(vol_jumpstart)

    def _take(self, m, j, qup, qlat):
        """Abstraction (+) / injection (-): storage first, then upstream inflow, then lateral flow (irf_route.f90:114-142 and
        the identical blocks of the Euler schemes).  Returns (inflow left, lateral flow left)."""
        if self.wm_flux is None:
            self.WMA[m][j] = 0.0
            return qup, qlat
        want = self.wm_flux[j]
        self.WMA[m][j] = want
        if want == -9999.0:
            return qup, qlat
        dt = self.o.dt
        if want <= 0:
            return qup, qlat - want
        if self.V1[m][j] / dt > want:
            self.V1[m][j] = self.V1[m][j] - want * dt
            return qup, qlat
        rest = want - self.V1[m][j] / dt
        self.V1[m][j] = 0.0
        if qup > rest:
            return qup - rest, qlat
        rest = rest - qup
        if qlat > rest:
            return 0.0, qlat - rest
        rest = rest - qlat
        self.WMA[m][j] = want - rest
        return 0.0, 0.0

    # -- step ---------------------------------------------------------------------------------
    def _basin2reach(self, flux):
        o, n = self.o, self.n
        rr = [0.0] * n
        for j in range(n):
            if self.hrus[j]:
                r = 0.0
                for w, hh in zip(self.wgt[j], self.hrus[j]):
                    ro = float(flux[hh])
                    if ro < -1.0e-3:
                        raise RouteError(20, "basin2reach/exceeded negative runoff tolerance")
                    r = r + w * ro * self.tc * self.lc
                if r < o.runoffMin:
                    r = o.runoffMin
                rr[j] = r * self.bas[j]
            else:
                rr[j] = o.runoffMin
        return rr

    def step(self, runoff, evapo=None, precip=None):
        o = self.o
        n = self.n
        if self.qmod == 1:
            self._read_obs()
        elif self.qmod != 0:
            raise RouteError(1, "main_route/Error: qmodOption invalid")
        rr = self._basin2reach(runoff)
        # lake evaporation / precipitation go through the same basin2reach (main_route.f90:174-199); None = exactly zero
        self.has_ep = bool(o.is_lake_sim and evapo is not None and precip is not None)
        if self.has_ep:
            self.evap = self._basin2reach(evapo)
            self.prec = self._basin2reach(precip)
        if o.doesBasinRoute == 1:
            nb = len(self.ff)
            for j in range(n):
                self.QI[j] = rr[j]
                self.QR0[j] = self.QR1[j]
                qf = self.qfut[j]
                if self.islake[j]:
                    qf[0] = qf[0] + 1.0 * rr[j]
                    for k in range(1, nb):
                        qf[k] = qf[k] + 0.0 * rr[j]
                else:
                    for k in range(nb):
                        qf[k] = qf[k] + self.ff[k] * rr[j]
                self.QR1[j] = qf[0]
                del qf[0]
                qf.append(0.0)
        else:
            for j in range(n):
                self.QR0[j] = self.QR1[j]
                self.QR1[j] = rr[j]
        for m in self.methods:
            for j in self.order:
                if self.islake[j] and m != 0:
                    self._lake(j, m)
                elif m == 0:
                    self._sum(j)
                elif m == 1:
                    self._irf(j)
                elif m == 2:
                    self._kwt(j, self.T0, self.T1)
                elif m == 4:
                    self._mc(j)
                else:
                    self._kw_dw(j, m)
        self.T0 = self.T1
        self.T1 = self.T0 + float(o.dt)
        self.itime += 1

    MONTHS = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]

    def _h06(self, j, m, v, qup):
        """Release of a Hanasaki-2006 reservoir holding volume v.  The monthly mean inflows H06_I_*, the release coefficient
        H06_E_rel_ini and the inflow memory are per REACH (RPARAM / RCHFLX), i.e. shared by the routing methods."""
        P, dt = self.net.lake_params, self.o.dt
        g = lambda name: float(P[name][j])
        month, day = self._month_day()
        if g("H06_I_mem_F") != 0.0:
            years = int(g("H06_I_mem_L"))
            n31, n30 = int(math.floor(years * 31 * 86400.0 / dt)), int(math.floor(years * 30 * 86400.0 / dt))
            nfeb = int(math.floor(years * (28 if self.o.calendar == "noleap" else 28.25) * 86400.0 / dt))
            if not hasattr(self, "h06_mem"):
                self.h06_mem = {}
            if j not in self.h06_mem:
                self.h06_mem[j] = [[g("H06_I_" + mo)] * n31 for mo in self.MONTHS]
            else:
                row = self.h06_mem[j][month - 1]
                row.insert(0, qup)
                row.pop()
            for k, mo in enumerate(self.MONTHS):
                if mo == "Nov":
                    continue                      # not updated by the reference
                n = nfeb if mo == "Feb" else (n30 if mo in ("Apr", "Jun", "Sep") else n31)
                tot = 0.0
                for x in self.h06_mem[j][k][:n]:
                    tot += x
                P["H06_I_" + mo][j] = tot / n
        inflow = [g("H06_I_" + mo) for mo in self.MONTHS]
        demand = [g("H06_D_" + mo) for mo in self.MONTHS]
        tot_i = tot_d = 0.0
        for x in inflow:
            tot_i += x
        for x in demand:
            tot_d += x
        i_year, d_year = tot_i / 12, tot_d / 12
        c = g("H06_Smax") / (i_year * 365 * 86400.0)
        start_month = 0
        for i in range(12):
            if i_year <= inflow[i]:
                start_month = i + 2
        if month == start_month and day == 1:
            P["H06_E_rel_ini"][j] = v / (g("H06_alpha") * g("H06_Smax"))
        if int(g("H06_purpose")) == 1:
            if g("H06_envfact") * i_year <= d_year:
                target = inflow[month - 1] * g("H06_c1") + i_year * g("H06_c2") * (demand[month - 1] / d_year)
            else:
                target = i_year + demand[month - 1] - d_year
        else:
            target = i_year
        q = self.Q[m][j]
        if c >= g("H06_c_compare"):
            q = target * g("H06_E_rel_ini")
        elif 0 <= c < g("H06_c_compare"):
            r = (c / g("H06_denominator")) ** g("H06_exponent")
            q = g("H06_E_rel_ini") * target * r + qup * (1 - r)
        dead = g("H06_Smax") * g("H06_frac_Sdead")
        if v < dead:
            q = max(q - (dead - v) / dt, 0.0)
        elif v > g("H06_Smax"):
            q = q + (v - g("H06_Smax")) / dt
        return q

    def _month_day(self):
        import datetime as _dt
        if not getattr(self.o, "sim_start", None):
            raise RouteError(20, "the lake model needs the simulation start datetime")
        y, mo, d, sec = self.o.sim_start
        days = int(math.floor((sec + (self.itime - 1) * float(self.o.dt) + 1e-6) / 86400.0))
        if self.o.calendar == "noleap":
            ml = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
            doy = (sum(ml[:mo - 1]) + (d - 1) + days) % 365
            k = 0
            while doy >= ml[k]:
                doy -= ml[k]
                k += 1
            return k + 1, doy + 1
        now = _dt.datetime(y, mo, d) + _dt.timedelta(days=days)
        return now.month, now.day

    def _day_of_year(self):
        """dayofyear of simDatetime(1) (datetime_data.f90:209-218), via Python's calendar for the standard calendar."""
        import datetime as _dt
        if not getattr(self.o, "sim_start", None):
            raise RouteError(20, "HYPE needs the simulation start datetime")
        y, mo, d, sec = self.o.sim_start
        elapsed = sec + (self.itime - 1) * float(self.o.dt)
        if self.o.calendar == "noleap":
            cum = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334]
            return (cum[mo - 1] + (d - 1) + int(math.floor((elapsed + 1e-6) / 86400.0))) % 365 + 1
        now = _dt.datetime(y, mo, d) + _dt.timedelta(days=int(math.floor((elapsed + 1e-6) / 86400.0)))
        return now.timetuple().tm_yday

    def _wb(self, m, j, qup, qlat, precip=0.0, evapo=0.0):
        dt = self.o.dt
        dvol = self.V1[m][j] - self.V0[m][j]
        took = self.WMA[m][j] if self.wm_flux is not None else 0.0
        self.WB[m][j] = dvol - (qup * dt + qlat * dt + precip + (-1.0 * took * dt) + (-1.0 * self.Q[m][j] * dt) + evapo)

    def _sum(self, j):
        q = self.QR1[j]
        if self.ups[j]:
            s = 0.0
            for u in self.ups[j]:
                s = s + self.Q[0][u]
            q = q + s
        self.Q[0][j] = q

    def _irf(self, j):
        m, dt = 1, self.o.dt
        self.V0[m][j] = self.V1[m][j]
        qup = 0.0
        if self.ngood[j] > 0:
            for u in self.ups[j][: self.ngood[j]]:
                qup = qup + self.Q[m][u]
            qlat = self.QR1[j]
        elif self.o.hw_drain_point == 1:
            qup = qup + self.QR1[j]
            qlat = 0.0
        else:
            qlat = self.QR1[j]
        self.INF[m][j] = qup
        q_in = qup
        qup, qlat = self._take(m, j, qup, qlat)
        qf, uh = self.qf_irf[j], self.uh[j]
        if self.length[j] > self.o.min_length_route:
            for k in range(len(uh)):
                qf[k] = qf[k] + uh[k] * qup
            qf[0] = min((max(0.0, self.V1[m][j]) / dt + qup) * _f32(0.999), qf[0])      # single-precision literal, irf_route.f90:245
            self.V1[m][j] = self.V1[m][j] - (qf[0] - qup) * dt
            self.Q[m][j] = qf[0] + qlat
            del qf[0]
            qf.append(0.0)
        else:
            for k in range(len(qf)):
                qf[k] = 0.0
            qf[0] = qup
            self.Q[m][j] = qf[0] + qlat
            self.V0[m][j] = 0.0
            self.V1[m][j] = 0.0
        self._finish(m, j, q_in, qlat)

    # -- Euler schemes: kwe_route.f90, dfw_route.f90, mc_route.f90 ---------------------------------
    def _inflow(self, j, m):
        self.V0[m][j] = self.V1[m][j]
        qup, head = 0.0, True
        if self.ngood[j] > 0:
            head = False
            for u in self.ups[j][: self.ngood[j]]:
                qup = qup + self.Q[m][u]
            qlat = self.QR1[j]
        elif self.o.hw_drain_point == 1:
            qup = qup + self.QR1[j]
            qlat = 0.0
        else:
            qlat = self.QR1[j]
        self.INF[m][j] = qup
        return qup, qlat, head

    def _geom(self, j):
        return self.width[j], self.zc[j], self.slope[j], self.man_n[j], self.zf[j], self.depth[j]

    def _stage(self, j, m):
        v = self.V1[m][j]
        self.FLOOD[m][j] = v - self.storage[j] if v > self.storage[j] else 0.0
        b, zc, _, _, zf, bd = self._geom(j)
        self.ELE[m][j] = hy_water_height(v / self.length[j], b, zc, zf, bd)

    def _dry(self, j, m):
        self.V0[m][j] = self.V1[m][j] = 0.0
        self.FLOOD[m][j] = self.ELE[m][j] = 0.0

    def _kw_dw(self, j, m):
        dt, L = self.o.dt, self.length[j]
        q_in, qlat, head = self._inflow(j, m)
        qup, qlat = self._take(m, j, q_in, qlat)
        mol = self.mol[m][j]
        nm = len(mol)
        if (not head) or self.o.hw_drain_point == 1:
            if L > self.o.min_length_route:
                b, zc, s, n, zf, bd = self._geom(j)
                qbar = abs((qup + mol[0] + mol[nm - 2]) / 3.0)
                y = hy_flow_depth(qbar, b, zc, s, n, zf, bd)
                ck = hy_celerity(qbar, y, b, zc, s, n, zf, bd)
                dk = hy_diffusivity(qbar, y, b, zc, s, n, zf, bd) if m == 5 else 0.0
                cur = solve_ade(L, mol, dt, qup, ck, dk)
                if abs(cur[nm - 2]) > 0.0:
                    red = min((max(0.0, self.V1[m][j]) + dt * qup) * 0.999 / (cur[nm - 2] * dt), 1.0)
                    for i in range(1, nm):
                        cur[i] = cur[i] * red
                self.V1[m][j] = self.V1[m][j] + (qup - cur[nm - 2]) * dt
                self._stage(j, m)
                self.Q[m][j] = cur[nm - 2] + qlat
                self.mol[m][j] = cur
            else:
                self.Q[m][j] = qup + qlat
                self.mol[m][j] = [0.0] * (nm - 1) + [self.Q[m][j]]
                self._dry(j, m)
        else:
            self.Q[m][j] = qlat
            self.mol[m][j] = [0.0] * (nm - 1) + [qlat]
            self._dry(j, m)
        self._finish(m, j, q_in, qlat)

    def _mc(self, j):
        m, dt, L = 4, self.o.dt, self.length[j]
        q_in, qlat, head = self._inflow(j, m)
        qup, qlat = self._take(m, j, q_in, qlat)
        q00, q01 = self.mol[m][j]
        q10 = q11 = 0.0
        if (not head) or self.o.hw_drain_point == 1:
            q10 = qup
            if L > self.o.min_length_route:
                b, zc, s, n, zf, bd = self._geom(j)
                qbar = (q00 + q10 + q01) / 3.0
                if qbar > 1.0e-50:
                    y = hy_flow_depth(abs(qbar), b, zc, s, n, zf, bd)
                    ck = hy_celerity(abs(qbar), y, b, zc, s, n, zf, bd)
                    nsub, dtsub = 1, dt
                    if ck * (dt / L) > 1.0:
                        nsub = int(math.ceil(dt / L * ck))
                        dtsub = dt / nsub
                    qin = [q00] + [q10] * nsub
                    qout = [q01] + [0.0] * nsub
                    for ix in range(1, nsub + 1):
                        qbar = (qin[ix] + qin[ix - 1] + qout[ix - 1]) / 3.0
                        if qbar > 1.0e-50:
                            y = hy_flow_depth(abs(qbar), b, zc, s, n, zf, bd)
                            tw = hy_btop(y, b, zc, zf, bd)
                            ck = hy_celerity(abs(qbar), y, b, zc, s, n, zf, bd)
                            x = 0.5 * (1.0 - qbar / (tw * s * ck * L))
                            cn = ck * dtsub / L
                            den = 1 - x + cn * (1 - 0.5)
                            c0 = (-x + cn * (1 - 0.5)) / den
                            c1 = (x + cn * 0.5) / den
                            c2 = (1 - x - cn * 0.5) / den
                            qout[ix] = max(0.0, c0 * qin[ix] + c1 * qin[ix - 1] + c2 * qout[ix - 1])
                    tot = 0.0
                    for v in qout[1:]:
                        tot = tot + v
                    q11 = tot / float(nsub)
                    if abs(q11) > 0.0:
                        q11 = q11 * min((self.V1[m][j] / dt + q10) * _f32(0.999) / q11, 1.0)     # single-precision literal, mc_route.f90:352
                    self.V1[m][j] = self.V1[m][j] + (q10 - q11) * dt
                    self._stage(j, m)
                    self.Q[m][j] = q11 + qlat
                else:
                    q11 = 0.0
                    self.Q[m][j] = q11 + qlat
                    self.V1[m][j] = self.V1[m][j] + (q10 - q11) * dt
                    self._stage(j, m)
            else:
                q11 = qup
                self.Q[m][j] = qup + qlat
                self._dry(j, m)
        else:
            self.Q[m][j] = qlat
            self._dry(j, m)
        self.mol[m][j] = [q10, q11]
        self._finish(m, j, q_in, qlat)

    def _lake(self, j, m):
        dt, net = self.o.dt, self.net
        qup = 0.0
        for u in self.ups[j]:
            qup = qup + self.Q[m][u]
        lt = self.ltype[j]
        lp = lambda name: float(net.lake_params[name][j])
        follows = "LakeTargVol" in net.lake_params and float(net.lake_params["LakeTargVol"][j]) != 0.0
        target = self.wm_vol[j] if self.wm_vol is not None else 0.0
        if self.itime == 1 and self.wm_jump and follows:
            self.V1[m][j] = target
        elif self.itime == 1:
            if lt == 0:
                self.V1[m][j] = float(net.D03_S0[j])
            elif lt == 1:
                self.V1[m][j] = float(net.D03_MaxStorage[j])
            elif lt == 2:
                self.V1[m][j] = lp("H06_Smax")
            elif lt == 3:
                self.V1[m][j] = (lp("HYP_E_emr") - lp("HYP_E_zero")) * lp("HYP_A_avg")
            else:
                raise RouteError(20, "lake type not restated")
        self.V0[m][j] = self.V1[m][j]
        v = self.V1[m][j] + qup * dt
        if self.o.LakeInputOption in (1, 2):
            v = v + self.QR1[j] * dt
        if self.o.LakeInputOption in (0, 2):
            pr = self.prec[j] if self.has_ep else 0.0
            ev = self.evap[j] if self.has_ep else 0.0
            v = v + pr * dt
            if v > ev * dt:
                v = v - ev * dt
            else:                       # the lake dries out: evaporation is cut to what was there, for every later method too
                if self.has_ep:
                    self.evap[j] = v / dt
                v = 0.0
        self.WMA[m][j] = self.wm_flux[j] if self.wm_flux is not None else 0.0
        if self.wm_flux is not None and self.wm_flux[j] != -9999.0:      # lake_route.f90:176-193
            f = self.wm_flux[j]
            if f <= 0 or f * dt <= v:
                v = v - f * dt
            else:
                self.WMA[m][j] = v / dt
                v = 0.0
        if follows:                                                       # lake_route.f90:196-203
            if v < target:
                q = 0.0
            else:
                q = (v - target) / dt
                v = target
        elif lt == 0:
            q = 0.0
        elif lt == 1:
            s0, smax = float(net.D03_S0[j]), float(net.D03_MaxStorage[j])
            if v - s0 > 0:
                q = float(net.D03_Coefficient[j]) * (v - s0) * ((v - s0) / (smax - s0)) ** float(net.D03_Power[j])
            else:
                q = 0.0
            q = q / 86400.0
            q = min(q, v / dt)
            v = v - q * dt
        elif lt == 2:                   # Hanasaki 2006, lake_route.f90:231-396 (without water-management demand)
            q = self._h06(j, m, v, qup)
            v = v - q * dt
        elif lt == 3:                   # HYPE, lake_route.f90:398-438
            ele = v / lp("HYP_A_avg") + lp("HYP_E_zero")
            doy = self._day_of_year()
            f_sin = max(0.0, 1 + lp("HYP_Qrate_amp") * math.sin(2 * 3.14159265359 * (doy + int(lp("HYP_Qrate_phs"))) / 365))
            f_lin = min(max((ele - lp("HYP_E_min")) / (lp("HYP_E_lim") - lp("HYP_E_min")), 0.0), 1.0)
            f_prim = 1 if lp("HYP_prim_F") != 0.0 else 0
            q_prim = f_sin * f_lin * f_prim * lp("HYP_Qrate_prim")
            q_spill = 0.0
            if ele > lp("HYP_E_emr"):
                q_spill = lp("HYP_Qrate_emr") * (ele - lp("HYP_E_emr")) ** lp("HYP_Erate_emr")
            q_sim = q_prim + q_spill if lp("HYP_Qsim_mode") != 0.0 else max(q_prim, q_spill)
            q = min(q_sim, max(0.0, (ele - lp("HYP_E_min")) * lp("HYP_A_avg")) / dt)
            v = v - q * dt
        else:
            raise RouteError(20, "lake type not restated")
        self.V1[m][j] = v
        self.Q[m][j] = q
        pe = (self.prec[j] * dt, -1.0 * self.evap[j] * dt) if self.has_ep else (0.0, 0.0)
        self._wb(m, j, qup, self.QR1[j], pe[0], pe[1])

    # -- KWT ------------------------------------------------------------------------------------
    def _qexmul(self, j, t0, t1):
        ups = self.ups[j]
        nupb = len(ups)
        nupr = sum(1 for u in ups if self.ngood[u] > 0)
        if nupb + nupr == 1:
            return [self.QR1[ups[0]] / self.width[j]], [t1]
        series, width, ctime = [], [], []
        for u in ups:
            series.append([Wave(self.QR0[u], t0, t0, True), Wave(self.QR1[u], t1, t1, True)])
            width.append(1.0)
            ctime.append(t1)
        imax = nupb
        for u in ups:
            if self.ngood[u] > 0:
                kw = self.KW[u]
                if kw is None:
                    raise RouteError(20, "qexmul_rch/KWAVE is not associated")
                ns = len(kw)
                nr = sum(1 for w in kw if w.RF)
                nq = min(nr + 1, ns)
                series.append([w.copy() for w in kw[:nq]])
                self.KW[u] = [w.copy() for w in kw[nr - 1: ns]]
                width.append(self.width[u])
                ctime.append(series[-1][1].TR)
                imax += nr - 1
        nups = len(series)
        mflg = [False] * nups
        itim = [1] * nups
        QD, TD = [], []
        jold, iold = None, None
        while True:
            jups = ctime.index(min(ctime))
            if jups == jold and itim[jups] == iold:
                raise RouteError(20, "qexmul_rch/stuck in the continuous do-loop")
            jold, iold = jups, itim[jups]
            if not mflg[jups]:
                if not series[jups][itim[jups]].RF:
                    mflg[jups] = True
                    ctime[jups] = HUGE
                else:
                    told = TD[-1] if TD else -HUGE
                    ct = ctime[jups]
                    if ct < told:
                        raise RouteError(30, "qexmul_rch/expect process in order of time")
                    if ct != told:
                        qagg = 0.0
                        for iu in range(nups):
                            s = series[iu]
                            iw = itim[iu]
                            sc = width[iu] / self.width[j]
                            if iu == jups:
                                sf = s[iw].QF * sc
                            else:
                                ib = iw
                                if s[ib].TR >= ct:
                                    ib = iw - 1
                                ie = ib + 1
                                if ib < 0 or ie >= len(s) or s[ie].TR < ct or s[ib].TR > ct:
                                    raise RouteError(40, "qexmul_rch/the times are not ordered as we assume")
                                sl = (s[ie].QF - s[ib].QF) / (s[ie].TR - s[ib].TR)
                                sf = (s[ib].QF + sl * (ct - s[ib].TR)) * sc
                            qagg = qagg + sf
                        if len(QD) + 1 > imax:
                            raise RouteError(60, "qexmul_rch/QD_TEMP bounds exceeded")
                        QD.append(qagg)
                        TD.append(ct)
                    if itim[jups] == len(series[jups]) - 1:
                        mflg[jups] = True
                        ctime[jups] = HUGE
                    else:
                        itim[jups] += 1
                        ctime[jups] = series[jups][itim[jups]].TR
            if all(mflg):
                break
        return QD, TD

    def _kwt(self, j, t0, t1):
        m = 2
        if self.ngood[j] == 0:
            self.INF[m][j] = 0.0
            self.Q[m][j] = self.QR1[j]
            self.KW[j] = [Wave(-9999.0, -9999.0, -9999.0, False)]
            return
        dt = t1 - t0
        # getusq_rch
        lake_up = None
        if self.o.is_lake_sim:
            for u in self.ups[j]:
                if self.islake[u]:
                    lake_up = u
            if lake_up is not None and len(self.ups[j]) > 1:
                raise RouteError(10, "getusq_rch/lake outlet reach should have one upstream lake")
        if lake_up is not None:
            QD, TD = [self.Q[m][lake_up] / self.width[j]], [t1]
        else:
            QD, TD = self._qexmul(j, t0, t1)
        if self.KW[j] is None:
            self.KW[j] = [Wave(QD[0], t0 - dt, t0, True)]
        own = self.KW[j]
        Q = [w.QF for w in own] + list(QD)
        T = [w.TI for w in own] + list(TD)
        X = [w.TR for w in own] + [-9999.0] * len(QD)
        if min(Q) < 0.0:
            raise RouteError(20, "kwt_rch/negative flow extracted from upstream reach")
        qup = 0.0
        for u in self.ups[j][: self.ngood[j]]:
            qup = qup + self.Q[m][u]
        self.INF[m][j] = qup
        if len(Q) > MAXQPAR:
            Q, T, X = remove_rch(Q, T, X)
        K = math.sqrt(self.slope[j]) / self.man_n[j]
        if self.wm_flux is not None and self.wm_flux[j] != -9999.0:
            # extract_from_rch (kwt_route.f90:351-455): the waves are scaled by the share of the time-step mean flow that is
            # added (take > 0, its own sign convention) or removed (take < 0); the exit times it sets are redone by kinwav_rch
            take = self.wm_flux[j]
            tot = interp_rch(T, Q, t0, t1) * self.width[j]
            if take > 0.0:
                Q = [Q[0]] + [q * (1.0 + take / tot) for q in Q[1:]]
            elif take < 0.0 and abs(take) < tot:
                Q = [Q[0]] + [q * (1.0 - abs(take) / tot) for q in Q[1:]]
            else:
                Q = [0.0] * len(Q)
        rq, rt, rx, rf = kinwav_rch(K, self.length[j], t0, t1, Q[1:], T[1:])
        nq2 = len(rq)
        Q = [Q[0]] + rq
        T = [T[0]] + rt
        X = [X[0]] + rx
        F = [True] + rf
        nr = sum(1 for f in F if f) - 1
        nn = nq2 - nr
        if nr + 1 > nq2:
            raise RouteError(21, "kwt_rch/no non-routed particle left")
        qnew = interp_rch(X[: nr + 2], Q[: nr + 2], t0, t1)
        self.Q[m][j] = qnew * self.width[j] + self.QR1[j]
        q_end = Q[nr] + ((Q[nr + 1] - Q[nr]) / (X[nr + 1] - X[nr])) * (t1 - X[nr])
        timei = T[nr] + ((T[nr + 1] - T[nr]) / (X[nr + 1] - X[nr])) * (t1 - X[nr])
        kw = [Wave(Q[i], T[i], X[i], F[i]) for i in range(nr + 1)]
        kw.append(Wave(q_end, timei, t1, True))
        kw += [Wave(Q[i], T[i], X[i], F[i]) for i in range(nr + 1, nq2 + 1)]
        if self.downId[j] <= 0 or (self.o.is_lake_sim and self.lakeinlet[j]):
            kw = kw[nr + 1:]
            assert len(kw) == nn + 1
        self.KW[j] = kw
