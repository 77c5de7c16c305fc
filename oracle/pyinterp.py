"""TEST INFRASTRUCTURE ONLY: a numpy restatement of the reference's dopri_interpolate
(/root/reference/R/dopri.R:577-642), operation for operation -- findInterval on the entry
start times (:602), theta = (t - t_entry) / h_entry, theta1 = 1 - theta (:603-604), and the
two Horner forms exactly as written (:617-639; every numpy operation is one individually
rounded IEEE operation, as R's are).  The GPU tests compare the product's interpolation
entry points with it bit for bit.  Not imported by the product."""
import numpy as np


def dopri_interpolate(history, n, t):
    """history: [n_entries, order * n + 2] (an entry per row: coefficients, t, h -- the
    transpose of R's `history` attribute); returns [len(t), n] like the R function."""
    h = np.asarray(history, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    nh = h.shape[0]
    it, ih = h.shape[1] - 2, h.shape[1] - 1
    order = (h.shape[1] - 2) // n
    if order not in (5, 8) or order * n + 2 != h.shape[1]:
        raise ValueError("Corrupt history object: incorrect number of rows")
    tr = (h[0, it], h[nh - 1, it] + h[nh - 1, ih])
    if t.min() < tr[0] or t.max() > tr[1]:
        raise ValueError("Time falls outside of range of known history [%s, %s]" % tr)
    idx = np.searchsorted(h[:, it], t, side="right") - 1  # findInterval
    theta = (t - h[idx, it]) / h[idx, ih]
    theta1 = 1.0 - theta
    ret = np.empty((t.shape[0], n))
    c = [h[idx, k * n:(k + 1) * n] for k in range(order)]  # c[k][:, i]: coefficient k of y_i
    for i in range(n):
        if order == 5:
            ret[:, i] = c[0][:, i] + theta * (
                c[1][:, i] + theta1 * (c[2][:, i] + theta * (c[3][:, i] + theta1 * c[4][:, i])))
        else:
            tmp = c[4][:, i] + theta * (c[5][:, i] + theta1 * (c[6][:, i] + theta * c[7][:, i]))
            ret[:, i] = c[0][:, i] + theta * (
                c[1][:, i] + theta1 * (c[2][:, i] + theta * (c[3][:, i] + theta1 * tmp)))
    return ret
