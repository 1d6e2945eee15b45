"""Extended-precision (x87 80-bit `np.longdouble`, 64-bit mantissa) restatement of the mapping algebra, used to show
that where the parity tests allow more than 1e-10 between the device and the double-precision oracle, the oracle itself
is no closer to the exact solution: both carry an error of ~cond(C) * 2^-53, only with different rounding.

The cluster sets and partition-of-unity weights are taken from the oracle (they are index / double-precision data that
both sides share bit for bit); everything that is sensitive to conditioning - basis-function evaluation, factorisation,
triangular solves, polynomial fit, evaluation - is redone in long double with Gaussian elimination and partial pivoting.
Test infrastructure only."""
import numpy as np

LD = np.longdouble


def basis_ld(kind, param, r):
    r = np.asarray(r, dtype=LD)
    if kind.startswith("compact-polynomial"):
        p = np.minimum(r / LD(param), LD(1))
        q = LD(1) - p
        return {"c0": q ** 2, "c2": q ** 4 * (4 * p + 1), "c4": q ** 6 * (35 * p ** 2 + 18 * p + 3),
                "c6": q ** 8 * (32 * p ** 3 + 25 * p ** 2 + 8 * p + 1),
                "c8": q ** 10 * (1287 * p ** 4 + 1350 * p ** 3 + 630 * p ** 2 + 150 * p + 15)}[kind[-2:]]
    if kind == "compact-tps-c2":  # impl/BasisFunctions.hpp:328-335
        p = np.minimum(r / LD(param), LD(1))
        v = 1 - 30 * p ** 2 - 10 * p ** 3 + 45 * p ** 4 - 6 * p ** 5 - p ** 3 * 60 * np.log(np.maximum(p, LD(1e-14)))
        return np.where(p >= 1, LD(0), v)
    if kind == "gaussian":  # shape parameter; cut off at sqrt(-ln 1e-9) / s (impl/BasisFunctions.hpp:229-272)
        s = LD(param)
        cut = np.sqrt(-np.log(LD(1e-9))) / s
        return np.where(r > cut, LD(0), np.exp(-(s * r) ** 2))
    if kind == "inverse-multiquadrics":
        return 1 / np.sqrt(LD(param) ** 2 + r * r)
    if kind == "multiquadrics":
        return np.sqrt(LD(param) ** 2 + r * r)
    if kind == "volume-splines":
        return np.abs(r)
    if kind == "thin-plate-splines":
        return np.log(np.maximum(r, LD(1e-14))) * r * r
    raise ValueError(kind)


def solve_ld(a, b):
    """a x = b by Gaussian elimination with partial pivoting in long double (b: matrix of right-hand sides)."""
    a = np.array(a, dtype=LD)
    b = np.array(b, dtype=LD).reshape(a.shape[0], -1)
    n = a.shape[0]
    for k in range(n):
        p = k + int(np.argmax(np.abs(a[k:, k])))
        if p != k:
            a[[k, p]] = a[[p, k]]
            b[[k, p]] = b[[p, k]]
        f = a[k + 1:, k] / a[k, k]
        a[k + 1:, k:] -= f[:, None] * a[k, k:]
        b[k + 1:] -= f[:, None] * b[k]
    x = np.zeros_like(b)
    for k in range(n - 1, -1, -1):
        x[k] = (b[k] - a[k, k + 1:] @ x[k + 1:]) / a[k, k]
    return x


def lstsq_ld(q, f):
    """min |q x - f| through the normal equations in long double (q has at most 4 well-scaled columns)."""
    q = np.asarray(q, dtype=LD)
    return solve_ld(q.T @ q, q.T @ np.asarray(f, dtype=LD).reshape(q.shape[0], -1))


def _dist(a, b):
    d = np.asarray(a, dtype=LD)[:, None, :] - np.asarray(b, dtype=LD)[None, :, :]
    return np.sqrt((d * d).sum(-1))


def _poly_matrix(x, axes):
    x = np.asarray(x, dtype=LD)
    return np.concatenate([np.ones((x.shape[0], 1), dtype=LD), x[:, axes]], axis=1)


def solve_system_ld(kind, param, polynomial, inp, out, values, conservative, axes):
    """One RadialBasisFctSolver (RadialBasisFctSolver.hpp:413-468) in long double; values: (rows, dims)."""
    C = basis_ld(kind, param, _dist(inp, inp))
    A = basis_ld(kind, param, _dist(out, inp))
    v = np.asarray(values, dtype=LD)
    if polynomial == "off":
        return solve_ld(C, A.T @ v) if conservative else A @ solve_ld(C, v)
    Q, V = _poly_matrix(inp, axes), _poly_matrix(out, axes)
    if conservative:  # mu = C^-1 A^T u; out = mu + Q (Q^T Q)^-1 (V^T u - Q^T mu)
        mu = solve_ld(C, A.T @ v)
        return mu + Q @ solve_ld(Q.T @ Q, V.T @ v - Q.T @ mu)
    beta = lstsq_ld(Q, v)
    return A @ solve_ld(C, v - Q @ beta) + V @ beta


def pum_map_ld(orc_mapping, kind, param, polynomial, inp, out, values, dims, conservative, dim):
    """PartitionOfUnityMapping::mapConsistent / mapConservative (PartitionOfUnityMapping.hpp:343-380) over the oracle's
    clusters in long double. `inp` / `out` are Mapping::input() / output(); returns float64 (rows * dims,)."""
    sin, sout = (out, inp) if conservative else (inp, out)  # solver-side meshes (PartitionOfUnityMapping.hpp:190-198)
    n_res = sin.shape[0] if conservative else sout.shape[0]
    res = np.zeros((n_res, dims), dtype=LD)
    v = np.asarray(values, dtype=LD).reshape(-1, dims)
    axes = list(range(dim))
    for c in range(orc_mapping.num_clusters):
        k = orc_mapping.cluster(c)
        ii, oo, w = k["in_ids"], k["out_ids"], np.asarray(k["weights"], dtype=LD)
        if len(oo) == 0 and not conservative:
            continue
        assert k["poly_params"] == 1 + dim or polynomial == "off", "the extended-precision model has no axis-drop branch"
        if conservative:
            res[ii] += solve_system_ld(kind, param, polynomial, sin[ii], sout[oo], v[oo] * w[:, None], True, axes)
        else:
            res[oo] += w[:, None] * solve_system_ld(kind, param, polynomial, sin[ii], sout[oo], v[ii], False, axes)
    return np.asarray(res, dtype=np.float64).reshape(-1)


def global_map_ld(kind, param, polynomial, inp, out, values, dims, conservative, dim):
    sin, sout = (out, inp) if conservative else (inp, out)
    v = np.asarray(values, dtype=LD).reshape(-1, dims)
    r = solve_system_ld(kind, param, polynomial, sin, sout, v, conservative, list(range(dim)))
    return np.asarray(r, dtype=np.float64).reshape(-1)
