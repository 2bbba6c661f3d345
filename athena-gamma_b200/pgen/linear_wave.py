"""Linear-wave convergence problem (src/pgen/linear_wave.cpp:63-190 setup, :430-575 pgen):
a sinusoidal eigenmode of amplitude `amp` on a uniform background, propagating obliquely
(angles from the box aspect ratio); MHD face fields come from a vector potential so that
div B = 0 to round-off.  Right eigenvectors are those of the Roe matrices the reference uses
(Eigensystem, linear_wave.cpp:626-...; same algebra as rsolvers/mhd/roe_mhd.cpp:245-470)."""
import numpy as np

from ._util import active, coords, empty_state


def right_eigenvector_mhd(wave, d, v1, v2, v3, h, b1, b2, b3, gm1, x=0.0, y=1.0):
    di = 1.0 / d
    btsq = b2 * b2 + b3 * b3
    vaxsq = b1 * b1 * di
    vsq = v1 * v1 + v2 * v2 + v3 * v3
    hp = h - (vaxsq + btsq * di)
    bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq
    twid_asq = max(gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x, 1e-20)
    ct2 = bt_starsq * di
    tsum, tdif = vaxsq + ct2 + twid_asq, vaxsq + ct2 - twid_asq
    cfsq = 0.5 * (tsum + np.sqrt(tdif * tdif + 4.0 * twid_asq * ct2))
    cf = np.sqrt(cfsq)
    cssq = twid_asq * vaxsq / cfsq
    cs = np.sqrt(cssq)
    bt, bt_star = np.sqrt(btsq), np.sqrt(bt_starsq)
    bet2, bet3 = (b2 / bt, b3 / bt) if bt != 0.0 else (1.0, 0.0)
    den = np.sqrt(gm1 - (gm1 - 1.0) * y)
    bet2s, bet3s = bet2 / den, bet3 / den
    bet_starsq = bet2s ** 2 + bet3s ** 2
    vbet = v2 * bet2s + v3 * bet3s
    if (cfsq - cssq) == 0.0:
        af_, as_ = 1.0, 0.0
    elif (twid_asq - cssq) <= 0.0:
        af_, as_ = 0.0, 1.0
    elif (cfsq - twid_asq) <= 0.0:
        af_, as_ = 1.0, 0.0
    else:
        af_ = np.sqrt((twid_asq - cssq) / (cfsq - cssq))
        as_ = np.sqrt((cfsq - twid_asq) / (cfsq - cssq))
    sqrtd = np.sqrt(d)
    isqrtd = 1.0 / sqrtd
    s = -1.0 if b1 < 0 else 1.0
    twid_a = np.sqrt(twid_asq)
    qf, qs = cf * af_ * s, cs * as_ * s
    afp, asp = twid_a * af_ * isqrtd, twid_a * as_ * isqrtd
    afpbb, aspbb = afp * bt_star * bet_starsq, asp * bt_star * bet_starsq
    vax = np.sqrt(vaxsq)
    ev = [v1 - cf, v1 - vax, v1 - cs, v1, v1 + cs, v1 + vax, v1 + cf]
    cols = {
        0: [af_, af_ * (v1 - cf), af_ * v2 + qs * bet2s, af_ * v3 + qs * bet3s,
            af_ * (hp - v1 * cf) + qs * vbet + aspbb, asp * bet2s, asp * bet3s],
        1: [0.0, 0.0, -bet3, bet2, -(v2 * bet3 - v3 * bet2), -bet3 * s * isqrtd, bet2 * s * isqrtd],
        2: [as_, as_ * (v1 - cs), as_ * v2 - qf * bet2s, as_ * v3 - qf * bet3s,
            as_ * (hp - v1 * cs) - qf * vbet - afpbb, -afp * bet2s, -afp * bet3s],
        3: [1.0, v1, v2, v3, 0.5 * vsq + (gm1 - 1.0) * x / gm1, 0.0, 0.0],
        4: [as_, as_ * (v1 + cs), as_ * v2 + qf * bet2s, as_ * v3 + qf * bet3s,
            as_ * (hp + v1 * cs) + qf * vbet - afpbb, -afp * bet2s, -afp * bet3s],
        5: [0.0, 0.0, bet3, -bet2, (v2 * bet3 - v3 * bet2), -bet3 * s * isqrtd, bet2 * s * isqrtd],
        6: [af_, af_ * (v1 + cf), af_ * v2 - qs * bet2s, af_ * v3 - qs * bet3s,
            af_ * (hp + v1 * cf) - qs * vbet + aspbb, asp * bet2s, asp * bet3s],
    }
    return np.array(cols[wave]), ev[wave]


def right_eigenvector_hydro(wave, v1, v2, v3, h, gm1):
    vsq = v1 * v1 + v2 * v2 + v3 * v3
    cs = np.sqrt(gm1 * max(h - 0.5 * vsq, 1e-20))
    cols = {0: [1.0, v1 - cs, v2, v3, h - v1 * cs], 1: [0.0, 0.0, 1.0, 0.0, v2],
            2: [0.0, 0.0, 0.0, 1.0, v3], 3: [1.0, v1, v2, v3, 0.5 * vsq],
            4: [1.0, v1 + cs, v2, v3, h + v1 * cs]}
    ev = [v1 - cs, v1, v1, v1, v1 + cs]
    return np.array(cols[wave]), ev[wave]


def setup(pin, mhd, f2, f3):
    """Mesh::InitUserMeshData (linear_wave.cpp:63-190)"""
    x1s = pin.get_real("mesh", "x1max") - pin.get_real("mesh", "x1min")
    x2s = pin.get_real("mesh", "x2max") - pin.get_real("mesh", "x2min")
    x3s = pin.get_real("mesh", "x3max") - pin.get_real("mesh", "x3min")
    ang_3 = pin.get_or_add_real("problem", "ang_3", -999.9)
    ang_2 = pin.get_or_add_real("problem", "ang_2", -999.9)
    if ang_3 == -999.9:
        ang_3 = np.arctan(x1s / x2s)
    sa3, ca3 = np.sin(ang_3), np.cos(ang_3)
    if pin.get_or_add_boolean("problem", "ang_3_vert", False):
        sa3, ca3, ang_3 = 1.0, 0.0, 0.5 * np.pi
    if ang_2 == -999.9:
        ang_2 = np.arctan(0.5 * (x1s * ca3 + x2s * sa3) / x3s)
    sa2, ca2 = np.sin(ang_2), np.cos(ang_2)
    if pin.get_or_add_boolean("problem", "ang_2_vert", False):
        sa2, ca2, ang_2 = 1.0, 0.0, 0.5 * np.pi
    lam = x1s * ca2 * ca3
    if f2 and ang_3 != 0.0:
        lam = min(lam, x2s * ca2 * sa3)
    if f3 and ang_2 != 0.0:
        lam = min(lam, x3s * sa2)
    return dict(sa2=sa2, ca2=ca2, sa3=sa3, ca3=ca3, k_par=2.0 * np.pi / lam)


def linear_wave(pmb, pin, out=None):
    m = pmb.pmy_mesh
    mhd = m.mhd
    f2, f3 = pmb.ncells2 > 1, pmb.ncells3 > 1
    g = setup(pin, mhd, f2, f3)
    sa2, ca2, sa3, ca3, k_par = g["sa2"], g["ca2"], g["sa3"], g["ca3"], g["k_par"]
    wave = pin.get_integer("problem", "wave_flag")
    amp = pin.get_real("problem", "amp")
    vflow = pin.get_or_add_real("problem", "vflow", 0.0)
    gam = pin.get_real("hydro", "gamma")
    gm1 = gam - 1.0
    d0, p0, u0 = 1.0, 1.0 / gam, vflow
    bx0, by0, bz0 = 1.0, np.sqrt(2.0), 0.5
    h0 = ((p0 / gm1 + 0.5 * d0 * u0 * u0) + p0) / d0
    if mhd:
        h0 += (bx0 ** 2 + by0 ** 2 + bz0 ** 2) / d0
        rem, _ = right_eigenvector_mhd(wave, d0, u0, 0.0, 0.0, h0, bx0, by0, bz0, gm1)
    else:
        rem, _ = right_eigenvector_hydro(wave, u0, 0.0, 0.0, h0, gm1)
    c = coords(pmb)
    out = empty_state(pmb, mhd, out)
    k, j, i = active(pmb)
    X = c["x1v"][i][None, None, :]
    Y = c["x2v"][j][None, :, None]
    Z = c["x3v"][k][:, None, None]
    x = ca2 * (X * ca3 + Y * sa3) + Z * sa2
    sn = np.sin(k_par * x)
    u = out["u"]
    u[0][k, j, i] = d0 + amp * sn * rem[0]
    mx = d0 * vflow + amp * sn * rem[1]
    my = amp * sn * rem[2]
    mz = amp * sn * rem[3]
    u[1][k, j, i] = mx * ca2 * ca3 - my * sa3 - mz * sa2 * ca3
    u[2][k, j, i] = mx * ca2 * sa3 + my * ca3 - mz * sa2 * sa3
    u[3][k, j, i] = mx * sa2 + mz * ca2
    u[4][k, j, i] = p0 / gm1 + 0.5 * d0 * u0 * u0 + amp * sn * rem[4]
    if mhd:
        u[4][k, j, i] += 0.5 * (bx0 ** 2 + by0 ** 2 + bz0 ** 2)
        dby, dbz = amp * rem[5], amp * rem[6]

        def pot(x1, x2, x3):
            xx = x1 * ca2 * ca3 + x2 * ca2 * sa3 + x3 * sa2
            yy = -x1 * sa3 + x2 * ca3
            Ay = bz0 * xx - (dbz / k_par) * np.cos(k_par * xx)
            Az = -by0 * xx + (dby / k_par) * np.cos(k_par * xx) + bx0 * yy
            return (-Ay * sa3 - Az * sa2 * ca3, Ay * ca3 - Az * sa2 * sa3, Az * ca2)

        x1f, x2f, x3f = c["x1f"], c["x2f"], c["x3f"]
        x1v, x2v, x3v = c["x1v"], c["x2v"], c["x3v"]
        n1, n2, n3 = pmb.ncells1, pmb.ncells2, pmb.ncells3

        def grid(a1, a2, a3):
            return np.meshgrid(a3, a2, a1, indexing="ij")

        # a1 at (x1v, x2f, x3f); a2 at (x1f, x2v, x3f); a3 at (x1f, x2f, x3v)
        x2fe = x2f if f2 else np.array([x2f[0], x2f[1]])
        x3fe = x3f if f3 else np.array([x3f[0], x3f[1]])
        Z3, Y2, X1 = grid(x1v, x2fe, x3fe)
        a1 = pot(X1, Y2, Z3)[0]
        Z3, Y2, X1 = grid(x1f, x2v, x3fe)
        a2 = pot(X1, Y2, Z3)[1]
        Z3, Y2, X1 = grid(x1f, x2fe, x3v)
        a3 = pot(X1, Y2, Z3)[2]
        ks, ke, js, je, is_, ie = pmb.ks, pmb.ke, pmb.js, pmb.je, pmb.is_, pmb.ie
        dx1 = c["dx1f"][None, None, is_:ie + 1]
        dx2 = c["dx2f"][None, js:je + 1, None]
        dx3 = c["dx3f"][ks:ke + 1, None, None]
        K, J, I = slice(ks, ke + 1), slice(js, je + 1), slice(is_, ie + 1)
        K1, J1, I1 = slice(ks + 1, ke + 2), slice(js + 1, je + 2), slice(is_ + 1, ie + 2)
        KF, JF, IF = slice(ks, ke + 2), slice(js, je + 2), slice(is_, ie + 2)
        out["b1"][K, J, IF] = ((a3[K, J1, IF] - a3[K, J, IF]) / dx2
                               - (a2[K1, J, IF] - a2[K, J, IF]) / dx3)
        out["b2"][K, JF, I] = ((a1[K1, JF, I] - a1[K, JF, I]) / dx3
                               - (a3[K, JF, I1] - a3[K, JF, I]) / dx1)
        out["b3"][KF, J, I] = ((a2[KF, J, I1] - a2[KF, J, I]) / dx1
                               - (a1[KF, J1, I] - a1[KF, J, I]) / dx2)
    return out


def linear_wave_errors(mesh, pin):
    """Mesh::UserWorkAfterLoop of the linear-wave problem (src/pgen/linear_wave.cpp:190-428):
    volume-weighted L1 error of the conserved variables (and of the cell-centred field) against
    the analytic eigenmode evaluated at the cell centres, normalised by the domain volume; RMS
    over the variables.  Host-side hook: downloads u and bcc of every local MeshBlock.
    Returns dict(rms, l1 = [d, M1, M2, M3, E, (B1c, B2c, B3c)], max = [...])."""
    mhd = mesh.mhd
    first = mesh.my_blocks[0]
    f2, f3 = first.ncells2 > 1, first.ncells3 > 1
    g = setup(pin, mhd, f2, f3)
    sa2, ca2, sa3, ca3, k_par = g["sa2"], g["ca2"], g["sa3"], g["ca3"], g["k_par"]
    wave = pin.get_integer("problem", "wave_flag")
    amp = pin.get_real("problem", "amp")
    vflow = pin.get_or_add_real("problem", "vflow", 0.0)
    gam = pin.get_real("hydro", "gamma")
    gm1 = gam - 1.0
    d0, p0, u0 = 1.0, 1.0 / gam, vflow
    bx0, by0, bz0 = 1.0, np.sqrt(2.0), 0.5
    h0 = ((p0 / gm1 + 0.5 * d0 * u0 * u0) + p0) / d0
    if mhd:
        h0 += (bx0 ** 2 + by0 ** 2 + bz0 ** 2) / d0
        rem, _ = right_eigenvector_mhd(wave, d0, u0, 0.0, 0.0, h0, bx0, by0, bz0, gm1)
    else:
        rem, _ = right_eigenvector_hydro(wave, u0, 0.0, 0.0, h0, gm1)
    nv = 8 if mhd else 5
    l1 = np.zeros(nv)
    mx_err = np.zeros(nv)
    for pmb in mesh.my_blocks:
        c = coords(pmb)
        k, j, i = active(pmb)
        X = c["x1v"][i][None, None, :]
        Y = c["x2v"][j][None, :, None]
        Z = c["x3v"][k][:, None, None]
        x = ca2 * (X * ca3 + Y * sa3) + Z * sa2
        sn = np.sin(k_par * x)
        mx = d0 * vflow + amp * sn * rem[1]
        my = amp * sn * rem[2]
        mz = amp * sn * rem[3]
        ana = [d0 + amp * sn * rem[0],
               mx * ca2 * ca3 - my * sa3 - mz * sa2 * ca3,
               mx * ca2 * sa3 + my * ca3 - mz * sa2 * sa3,
               mx * sa2 + mz * ca2]
        e0 = p0 / gm1 + 0.5 * d0 * u0 * u0 + amp * sn * rem[4]
        if mhd:
            e0 = e0 + 0.5 * (bx0 * bx0 + by0 * by0 + bz0 * bz0)
            bx = bx0
            by = by0 + amp * sn * rem[5]
            bz = bz0 + amp * sn * rem[6]
            ana_b = [bx * ca2 * ca3 - by * sa3 - bz * sa2 * ca3,
                     bx * ca2 * sa3 + by * ca3 - bz * sa2 * sa3,
                     bx * sa2 + bz * ca2]
        ana.append(e0)
        vol = (c["dx1f"][i][None, None, :] * c["dx2f"][j][None, :, None]
               * c["dx3f"][k][:, None, None])
        u = pmb.get("u")
        for n in range(5):
            d = np.abs(ana[n] - u[n][k, j, i])
            l1[n] += float(np.sum(d * vol))
            mx_err[n] = max(mx_err[n], float(d.max()))
        if mhd:
            bcc = pmb.get("bcc")
            for n in range(3):
                d = np.abs(ana_b[n] - bcc[n][k, j, i])
                l1[5 + n] += float(np.sum(d * vol))
                mx_err[5 + n] = max(mx_err[5 + n], float(d.max()))
    p = mesh.params
    vol_mesh = (p.x1max - p.x1min) * (p.x2max - p.x2min) * (p.x3max - p.x3min)
    l1 /= vol_mesh
    return {"rms": float(np.sqrt(np.sum(l1 ** 2))), "l1": l1.tolist(), "max": mx_err.tolist()}
