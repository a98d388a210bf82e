"""Shared helpers for the parity tests."""
import numpy as np

KEYS = ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")
RTOL_F64 = 1e-12   # BASELINE.json north_star: FP64 relative error <= 1e-12


def fp32_tolerances(nodes, dx=0.7, cond=4.0, slack=8.0):
    """Per-array FP32 tolerances = eps32 x conditioning (R = max|x| / dx): Jacobians eps32 R, inverse / face Jacobians
    eps32 R cond(F), conserved metrics eps32 R^2 (products of a coordinate difference and a coordinate value), with a
    stated factor of slack.  The reference only asserts a GCL bound in Float32 (test_unified_grid_backends.jl:58-123)."""
    eps32 = float(np.finfo(np.float32).eps)
    R = max(float(np.abs(np.asarray(a)).max()) for a in nodes) / dx
    return {"cell_fwd": slack * eps32 * R, "cell_inv": slack * eps32 * R * cond, "face_fwd": slack * eps32 * R * cond,
            "face_inv": slack * eps32 * R * cond, "face_cons": slack * eps32 * R * R}


def assert_fp32_close(got, ref, nodes, where, label="", **kw):
    tol = fp32_tolerances(nodes, **kw)
    for k in KEYS:
        e = metric_err(np.asarray(got[k])[..., where, :], np.asarray(ref[k])[..., where, :], k != "face_cons")
        assert e <= tol[k], f"{label} {k}: rel err {e:.3e} > {tol[k]:.1e} (eps32 x conditioning)"



def metric_err(a, b, has_j):
    """Relative error per cell: matrix entries against the largest entry of the
    same matrix, J against itself.  Returns max over cells."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    K = a.shape[-1] - (1 if has_j else 0)
    ma, mb = a[..., :K], b[..., :K]
    scale = np.abs(mb).max(axis=-1, keepdims=True)
    scale = np.where(scale == 0, 1.0, scale)
    e = (np.abs(ma - mb) / scale).max() if ma.size else 0.0
    if has_j:
        ja, jb = a[..., K], b[..., K]
        e = max(e, (np.abs(ja - jb) / np.where(jb == 0, 1.0, np.abs(jb))).max())
    return float(e)


def assert_metrics_close(got, ref, rtol=RTOL_F64, keys=KEYS, where=None, label=""):
    """got/ref: dicts with packed arrays ([ncell,K] / [N,ncell,K]).  `where`: optional
    index array of cells to compare (e.g. cell.domain)."""
    for k in keys:
        if ref.get(k) is None or got.get(k) is None:
            continue
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        if where is not None:
            a, b = a[..., where, :], b[..., where, :]
        e = metric_err(a, b, has_j=k != "face_cons")
        assert e <= rtol, f"{label} {k}: rel err {e:.3e} > {rtol:.1e}"


def domain_indices(nfull, nhalo, shrink_axis=None):
    """Linear (Fortran) indices of cell.domain; optionally expand(domain, axis, -1)."""
    rngs = []
    for d, n in enumerate(nfull):
        lo, hi = nhalo, n - nhalo
        if shrink_axis is not None and d == shrink_axis:
            lo, hi = lo + 1, hi - 1
        rngs.append(np.arange(lo, hi))
    mesh = np.meshgrid(*rngs, indexing="ij")
    lin = np.zeros_like(mesh[0])
    stride = 1
    for d, n in enumerate(nfull):
        lin = lin + mesh[d] * stride
        stride *= n
    return lin.ravel(order="F")


def grid_outputs(cg, grid):
    """Copy a grid's FULL-profile outputs to packed numpy arrays like the oracle's."""
    N = grid.dim
    cm = cg.cell_metrics(grid)
    fm = cg.face_metrics(grid)
    KM = N * N + 1
    out = {
        "cell_fwd": cm.forward.reshape(-1, KM).cpu().numpy(),
        "cell_inv": cm.inverse.reshape(-1, KM).cpu().numpy(),
        "face_fwd": np.stack([f.forward.reshape(-1, KM).cpu().numpy() for f in fm]),
        "face_inv": np.stack([f.inverse.reshape(-1, KM).cpu().numpy() for f in fm]),
        "face_cons": np.stack([f.conserved.reshape(-1, N * N).cpu().numpy() for f in fm]),
        "nfull": grid.nfull,
    }
    return out


def mild_nodes(shape, seed=0, amp=0.08, noise=0.01):
    """Well-conditioned curvilinear test mesh: unit-spaced lattice + smooth waves +
    seeded noise (SURVEY.md §8d robustness set, numpy.random.default_rng(seed))."""
    rng = np.random.default_rng(seed)
    dim = len(shape)
    idx = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    out = []
    for c in range(dim):
        q = (0.7 + 0.2 * c) * idx[c]
        for d in range(dim):
            if d != c:
                q = q + amp * np.sin(0.9 * idx[d] + 0.3 * c) * np.cos(0.5 * idx[(d + 1) % dim])
        q = q + 0.05 * idx[(c + 1) % dim]  # shear
        q = q + noise * rng.standard_normal(shape)
        out.append(np.asfortranarray(q))
    return out


def sampled_outputs(cg, grid, sample):
    """FULL-profile outputs of the cells `sample` (linear Fortran indices of the window), gathered on the device."""
    import torch
    idx = torch.from_numpy(np.ascontiguousarray(sample)).to(f"cuda:{grid.device}")
    N = grid.dim
    K = N * N + 1
    cm, fm = cg.cell_metrics(grid), cg.face_metrics(grid)
    pick = lambda t, k: t.reshape(-1, k)[idx].cpu().numpy()
    return {"cell_fwd": pick(cm.forward, K), "cell_inv": pick(cm.inverse, K),
            "face_fwd": np.stack([pick(fm[a].forward, K) for a in range(N)]),
            "face_inv": np.stack([pick(fm[a].inverse, K) for a in range(N)]),
            "face_cons": np.stack([pick(fm[a].conserved, N * N) for a in range(N)])}


def roundoff_report(oracle, got, nodes, nhalo, scheme, sample=None, precision="ld"):
    """Three-way comparison per output array: CUDA (`got`) and the FP64 oracle, each against the same closure graph
    evaluated in extended precision ("truth"), and against each other.  Returns {key: {gpu_vs_truth,
    oracle64_vs_truth, gpu_vs_oracle64}} (metric_err: relative to the largest entry of the same matrix)."""
    truth = oracle.discrete_eval(nodes, nhalo, scheme, sample=sample, precision=precision)
    o64 = oracle.discrete_eval(nodes, nhalo, scheme, sample=sample)
    rep = {}
    for k in KEYS:
        hj = k != "face_cons"
        rep[k] = {"gpu_vs_truth": metric_err(got[k], truth[k], hj), "oracle64_vs_truth": metric_err(o64[k], truth[k], hj),
                  "gpu_vs_oracle64": metric_err(got[k], o64[k], hj)}
    return rep


def gcl_max_numpy(cons, nfull, nhalo, dtype=np.float64):
    """gcl.jl:20-117 in numpy with the accumulation order and type of k_gcl: per column c,
    acc = sum over axes (in order) of (C_ax[I][ax, c] - C_ax[I - e_ax][ax, c]) over cell.domain; max |acc|.
    cons: [N, ncell (Fortran order), N*N]."""
    N = len(nfull)
    dom = domain_indices(nfull, nhalo)
    strides = [int(np.prod(nfull[:d])) for d in range(N)]
    c = np.asarray(cons).astype(dtype)
    out = np.zeros(N)
    for col in range(N):
        acc = np.zeros(len(dom), dtype=dtype)
        for ax in range(N):
            e = ax + N * col
            acc = (acc + (c[ax][dom, e] - c[ax][dom - strides[ax], e])).astype(dtype)
        out[col] = float(np.abs(acc).max())
    return out


def is_domain(sample, nfull, nhalo):
    """Boolean mask: which linear (Fortran) cell indices of cell.full lie in cell.domain."""
    s = np.asarray(sample, dtype=np.int64)
    ok = np.ones(len(s), dtype=bool)
    for n in nfull:
        c = s % n
        s = s // n
        ok &= (c >= nhalo) & (c < n - nhalo)
    return ok


def assert_three_way(oracle, got, nodes, nhalo, scheme, sample, nfull, label="", domain_tol=1e-12):
    """The parity statement at BASELINE sizes.  `got`: CUDA outputs of the cells `sample`.
      * cell.domain (where parity is defined): every array within `domain_tol` = 1e-12 of the exact
        value of the reference's algorithm on these inputs (the oracle's closure graph evaluated in long double);
      * halo cells: their nodes are Line() extrapolations ROUNDED to the working precision (absolute error eps |x|), so
        both FP64 evaluations carry eps (|x| / dx)^2 there: the CUDA error is held to 2 x that of the FP64 oracle,
        or to the conditioning bound;
      * the FP64 oracle's own distance from the truth is returned for the record (it is the larger one on
        cell.domain at 512^3: 7e-12 against 6e-14)."""
    truth = oracle.discrete_eval(nodes, nhalo, scheme, sample=sample, precision="ld")
    o64 = oracle.discrete_eval(nodes, nhalo, scheme, sample=sample)
    dom = is_domain(sample, nfull, nhalo)
    rep = {}
    R = max(float(np.abs(np.asarray(a)).max()) for a in nodes) / min(
        float(np.abs(np.diff(np.asarray(a), axis=d)).max()) for d, a in enumerate(nodes))
    halo_bound = 8 * np.finfo(np.float64).eps * R * R
    for k in KEYS:
        hj = k != "face_cons"
        g, t, o = np.asarray(got[k]), truth[k], o64[k]
        rep[k] = {}
        if dom.any():
            e = metric_err(g[..., dom, :], t[..., dom, :], hj)
            rep[k]["domain_gpu_vs_truth"] = e
            rep[k]["domain_oracle64_vs_truth"] = metric_err(o[..., dom, :], t[..., dom, :], hj)
            assert e <= domain_tol, f"{label} {k} on cell.domain: CUDA vs long-double truth {e:.3e} > {domain_tol:.0e}"
        if (~dom).any():
            e = metric_err(g[..., ~dom, :], t[..., ~dom, :], hj)
            eo = metric_err(o[..., ~dom, :], t[..., ~dom, :], hj)
            rep[k]["halo_gpu_vs_truth"], rep[k]["halo_oracle64_vs_truth"] = e, eo
            assert e <= max(domain_tol, 2 * eo, halo_bound), f"{label} {k} in the halo: CUDA {e:.3e}, FP64 oracle {eo:.3e}, bound {halo_bound:.1e}"
    return rep
