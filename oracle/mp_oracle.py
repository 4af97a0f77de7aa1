"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

50-digit mpmath forward evaluation of the gp_grid log-density (default options: Matern/RBF/RQ kernels,
heteroscedastic or homoscedastic Normal likelihood, est_la on/off, no PC prior, not robust) following
inst/stan/gp_grid.stan:64-233 literally.  Used to tell apart "two FP64 implementations disagree" from
"one of them is wrong" at ill-conditioned points.  Only tests/ may import this.
"""
import mpmath as mp

mp.mp.dps = 50


def log_prob_mp(sd, u, jacobian=True):
    assert not sd.robust and not sd.pcprior
    u = [mp.mpf(float(v)) for v in u]
    n1, n2 = sd.n1, sd.n2
    k = 0
    lp = mp.mpf(0)

    def lub(x):
        nonlocal lp
        th = 1 / (1 + mp.e ** (-x))
        if jacobian:
            lp += mp.log(th) + mp.log(1 - th)
        return th

    def lb(x):
        nonlocal lp
        if jacobian:
            lp += x
        return mp.e ** x

    la1 = la2 = mp.mpf(0)
    if sd.est_la:
        la1 = lub(u[k]); k += 1
        la2 = lub(u[k]); k += 1
    m1, m2, th1, th2 = u[k], u[k + 1], u[k + 2], u[k + 3]; k += 4
    sl1 = lb(u[k]); k += 1
    sl2 = lb(u[k]); k += 1
    b1 = lb(u[k]); k += 1
    b2 = lb(u[k]); k += 1
    ell = lb(u[k]); k += 1
    sigf = lb(u[k]); k += 1
    alpha = None
    if sd.est_alpha:
        alpha = lb(u[k]); k += 1
    z = mp.matrix(n2, n1)
    for j in range(n1):
        for i in range(n2):
            z[i, j] = u[k]; k += 1
    s = lb(u[k]); k += 1
    s21 = lb(u[k]); k += 1
    s22 = lb(u[k]); k += 1
    x1 = [mp.mpf(float(v)) for v in sd.x1]
    x2 = [mp.mpf(float(v)) for v in sd.x2]

    def kern(d):
        if sd.kernel == 1:
            return mp.e ** (-(d * d) / (2 * ell ** 2))
        if sd.kernel == 2:
            if sd.nu_matern == 1:
                return mp.e ** (-d / ell)
            if sd.nu_matern == 2:
                return (1 + mp.sqrt(3) * d / ell) * mp.e ** (-mp.sqrt(3) * d / ell)
            return (1 + mp.sqrt(5) * d / ell + mp.mpf(5) / 3 * d * d / ell ** 2) * mp.e ** (-mp.sqrt(5) * d / ell)
        return mp.e ** (-alpha * mp.log(1 + d * d / (2 * alpha * ell ** 2)))

    c1 = mp.matrix(n1, n1)
    c2 = mp.matrix(n2, n2)
    jit = mp.mpf("1e-10")
    for a in range(n1):
        for b in range(n1):
            c1[a, b] = sigf * kern(abs(x1[a] - x1[b])) + (jit if a == b else 0)
    for a in range(n2):
        for b in range(n2):
            c2[a, b] = kern(abs(x2[a] - x2[b])) + (jit if a == b else 0)
    L1 = mp.cholesky(c1)
    L2 = mp.cholesky(c2)
    GP = L2 * z * L1.T
    ten = mp.mpf(10)
    p01 = [la1 + (1 - la1) / (1 + ten ** (sl1 * (x1[j] - m1))) for j in range(n1)]
    p02 = [la2 + (1 - la2) / (1 + ten ** (sl2 * (x2[i] - m2))) for i in range(n2)]
    f = mp.matrix(n2 + 1, n1 + 1)
    f[0, 0] = 1
    for j in range(n1):
        f[0, j + 1] = p01[j]
    for i in range(n2):
        f[i + 1, 0] = p02[i]
    for j in range(n1):
        for i in range(n2):
            p0 = p01[j] * p02[i]
            lg = mp.log(p0 / (1 - p0))
            B = GP[i, j]
            D = -p0 / (1 + mp.e ** (b1 * B + lg)) + (1 - p0) / (1 + mp.e ** (-b2 * B - lg))
            f[i + 1, j + 1] = p0 + D
    h2pi = mp.log(2 * mp.pi) / 2

    def normal(y, mu, sd_):
        return -h2pi - mp.log(sd_) - ((y - mu) / sd_) ** 2 / 2

    def invgamma(y, a, b):
        return a * mp.log(b) - mp.loggamma(a) - (a + 1) * mp.log(y) - b / y

    lp += -mp.log(mp.pi) - mp.log(1 + s * s) - mp.log(mp.mpf(1) / 2)
    lp += invgamma(s21, 3, 2) + invgamma(s22, 3, 2)
    if sd.est_la:
        cb = mp.loggamma(mp.mpf("2.25")) - mp.loggamma(1) - mp.loggamma(mp.mpf("1.25"))
        lp += cb + mp.mpf("0.25") * mp.log(1 - la1) + cb + mp.mpf("0.25") * mp.log(1 - la2)
    lp += -sl1 - sl2
    lp += normal(th1, 0, 1) + normal(th2, 0, 1)
    lp += normal(m1, th1, mp.sqrt(s21)) + normal(m2, th2, mp.sqrt(s22))
    lp += normal(b1, 1, mp.mpf("0.1")) + normal(b2, 1, mp.mpf("0.1"))
    lp += invgamma(ell, 5, 5)
    lp += -h2pi - mp.log(sigf) - (mp.log(sigf) - 1) ** 2 / 2
    if sd.est_alpha:
        lp += -alpha
    for j in range(n1):
        for i in range(n2):
            lp += normal(z[i, j], 0, 1)
    lam = mp.mpf(float(sd.lambda_))
    ld = n2 + 1
    for n in range(len(sd.y)):
        c = int(sd.ii_obs[n]) - 1
        fv = f[c % ld, c // ld]
        sdv = s * mp.sqrt(fv + lam) if sd.heteroscedastic else s
        lp += normal(mp.mpf(float(sd.y[n])), fv, sdv)
    return lp
