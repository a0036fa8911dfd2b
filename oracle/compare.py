"""Nested-result comparison helpers shared by the parity tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import numpy as np


def _num(v):
    if isinstance(v, str):
        return {"inf": np.inf, "-inf": -np.inf, "nan": np.nan}[v]
    return v


def compare_results(ref, got, rtol, atol, path="", problems=None, int_keys=("time_dce", "max_obst_risk_index")):
    """Recursively compare a reference result dict (possibly JSON-decoded: str keys, 'inf' strings)
    with a computed one.  Returns a list of human-readable mismatches (empty = equal)."""
    problems = [] if problems is None else problems
    if isinstance(ref, dict):
        gk = {str(k): k for k in got.keys()} if isinstance(got, dict) else {}
        if not isinstance(got, dict) or set(map(str, ref.keys())) != set(gk.keys()):
            problems.append(f"{path}: key sets differ: {sorted(map(str, ref.keys()))} vs "
                            f"{sorted(gk.keys()) if isinstance(got, dict) else type(got)}")
            return problems
        for k, v in ref.items():
            compare_results(v, got[gk[str(k)]], rtol, atol, f"{path}/{k}", problems, int_keys)
        return problems
    if isinstance(ref, (list, tuple, np.ndarray)):
        r = np.array([_num(x) for x in ref], dtype=np.float64)
        g = np.asarray(got, dtype=np.float64)
        if r.shape != g.shape:
            problems.append(f"{path}: shape {r.shape} vs {g.shape}")
            return problems
        bad = ~np.isclose(g, r, rtol=rtol, atol=atol, equal_nan=True)
        if bad.any():
            i = int(np.argmax(bad))
            problems.append(f"{path}[{i}]: ref {r[i]!r} got {g[i]!r} ({int(bad.sum())} of {bad.size} differ)")
        return problems
    r = _num(ref)
    key = path.rsplit("/", 1)[-1]
    if key in int_keys or isinstance(r, (bool, np.bool_)):
        if int(r) != int(got):
            problems.append(f"{path}: ref {r!r} got {got!r} (exact)")
        return problems
    r, g = float(r), float(got)
    if not np.isclose(g, r, rtol=rtol, atol=atol, equal_nan=True):
        problems.append(f"{path}: ref {r!r} got {g!r}")
    return problems
