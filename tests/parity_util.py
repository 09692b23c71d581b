"""Helpers shared by the parity tests: which positions / bank columns of two runs are comparable when a few knife-edge
arg-max rows differ, and a record of how much was (not) compared so that lost coverage stays visible.

Images are independent (models/IPSRFunction.py:46) and inside an image
  * an UNMASKED output position q depends on ind[q] only                                  (:129-131);
  * the masked position of step l depends on the matches of the masked steps 0..l         (:93-126);
  * the gradient of bank column p collects the rows routed to p (:156-173): unit routes of unmasked rows (and step 0)
    and, for blended rows, entries that survive the int64 store -- those move with every weight of the chain.
"""
import numpy as np

COVERAGE = {"cases": 0, "positions": 0, "positions_compared": 0, "columns": 0, "columns_compared": 0,
            "rows_differing": 0, "rows_unsafe": 0, "cases_with_differences": 0}


def comparable(ind_a, ind_b, flag):
    """ind_a, ind_b [B,N]; flag [N] or [B,N].  Returns (pos_ok [B,N], col_ok [B,N])."""
    ind_a, ind_b = np.asarray(ind_a), np.asarray(ind_b)
    B, N = ind_a.shape
    flag = np.broadcast_to(np.asarray(flag).reshape(-1, N), (B, N))
    same = ind_a == ind_b
    pos_ok = same.copy()
    col_ok = np.ones((B, N), bool)
    for b in range(B):
        midx = np.nonzero(flag[b])[0]
        diverged = False
        if len(midx):
            chain_same = np.logical_and.accumulate(same[b, midx])
            pos_ok[b, midx] = chain_same
            diverged = not chain_same[-1]
        for q in np.nonzero(~same[b])[0]:
            col_ok[b, ind_a[b, q]] = False
            col_ok[b, ind_b[b, q]] = False
        if diverged:                                   # the blend weights differ from the first differing step on
            col_ok[b, ind_a[b, midx]] = False
            col_ok[b, ind_b[b, midx]] = False
    return pos_ok, col_ok


def record(pos_ok, col_ok, ind_a, ind_b, safe):
    """Book-keeping + the bound the parity bar allows: rows may differ only where the fp64 top-2 gap is <= 1e-4."""
    ndiff = int((np.asarray(ind_a) != np.asarray(ind_b)).sum())
    nunsafe = int((~np.asarray(safe)).sum())
    COVERAGE["cases"] += 1
    COVERAGE["positions"] += pos_ok.size
    COVERAGE["positions_compared"] += int(pos_ok.sum())
    COVERAGE["columns"] += col_ok.size
    COVERAGE["columns_compared"] += int(col_ok.sum())
    COVERAGE["rows_differing"] += ndiff
    COVERAGE["rows_unsafe"] += nunsafe
    COVERAGE["cases_with_differences"] += int(ndiff > 0)
    assert ndiff <= nunsafe, "%d arg-max rows differ but only %d rows have an fp64 top-2 gap <= 1e-4" % (ndiff, nunsafe)
    assert not (np.asarray(ind_a) != np.asarray(ind_b))[np.asarray(safe)].any()
    return ndiff


def expand(mask_bn, C, H, W):
    """[B,N] position / column mask -> [B,C,H,W]."""
    B, N = mask_bn.shape
    return np.broadcast_to(mask_bn[:, None, :], (B, C, N)).reshape(B, C, H, W)
