"""Monte-Carlo statistics of a detector bin from the sums Sum_i w_i^k, k = 0..4, that FluxRecorder records per bin
(FluxRecorder.hpp:50-63; the MCNP manual and Camps & Baes 2018, which the reference cites): the relative error R, the
variance of the variance VOV, and the reference's own reliability rule (R < 0.1 and VOV < 0.1).  Rows are
(N, Sum w, Sum w^2, Sum w^3, Sum w^4), one column per bin.  FluxRecorder.hpp:50-63 defines N as the number of packets LAUNCHED
during the peel-off segments (w_i = 0 for a history that does not reach the bin); row 0 holds the number of histories that
did reach it.  `launched` gives that N; without it the functions fall back on row 0, which underestimates R where only a small
part of the histories reaches a bin (many wavelength bins)."""
import numpy as np


def _with_launched(stats, launched):
    if launched is None:
        return stats
    out = [np.asarray(stats[k], dtype=float) for k in range(5)]
    out[0] = np.full_like(out[1], float(launched))
    return out


def rel_error(stats, launched=None):
    stats = _with_launched(stats, launched)
    n, w1, w2 = stats[0], stats[1], stats[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.maximum(w2 / (w1 * w1) - 1.0 / np.maximum(n, 1), 0.0))


def vov(stats, launched=None):
    """Variance of the variance: Sum (w - mean)^4 / (Sum (w - mean)^2)^2 - 1/N."""
    stats = _with_launched(stats, launched)
    n, s1, s2, s3, s4 = (np.asarray(stats[k], dtype=float) for k in range(5))
    with np.errstate(divide="ignore", invalid="ignore"):
        m2 = s2 - s1 * s1 / n
        m4 = s4 - 4 * s1 * s3 / n + 6 * s1 * s1 * s2 / n ** 2 - 3 * s1 ** 4 / n ** 3
        v = m4 / (m2 * m2) - 1.0 / n
    return np.where(np.isfinite(v), v, np.inf)


def reliable(stats, rmax=0.1, vovmax=0.1, launched=None):
    """The bins whose error estimate can be trusted: R < 0.1 and VOV < 0.1 (Camps & Baes 2018, section 3.3)."""
    r = rel_error(stats, launched)
    return np.isfinite(r) & (r < rmax) & (vov(stats, launched) < vovmax) & (np.asarray(stats[1]) > 0)
