"""Adhesion outlines from the device's marching-squares cell records.

`ipb_fa_contour_cells` emits, for every 2 x 2 cell of a crop's label map and every label cut by it,
{cell index, label << 4 | case}.  For a binary mask at level 0.5 (the reference's call,
INT/FA_Analyzer.py:168: find_contours(labeled_img == k, 0.5)) every contour point is the midpoint of
a pixel pair, so these records hold everything skimage's _get_contour_segments produces.  Here the
records of one label are put back into raster order, expanded to oriented segments by the case
table, and linked with the algorithm of skimage's _assemble_contours, so the polylines -- point
values, point order, contour order -- are the reference's.  The linking is O(records) per label;
the O(adhesions x crop pixels) scan of the reference is gone.
"""
from collections import deque

import numpy as np

# case -> oriented segments as (from side, to side); sides: 0 top, 1 bottom, 2 left, 3 right
_T, _B, _L, _R = 0, 1, 2, 3
CASE_SEGMENTS = {1: ((_T, _L),), 2: ((_R, _T),), 3: ((_R, _L),), 4: ((_L, _B),), 5: ((_T, _B),),
                 6: ((_R, _T), (_L, _B)), 7: ((_R, _B),), 8: ((_B, _R),), 9: ((_T, _L), (_B, _R)),
                 10: ((_B, _T),), 11: ((_B, _L),), 12: ((_L, _R),), 13: ((_T, _R),), 14: ((_L, _T),)}


def _side_point(r0, c0, side):
    # twice the coordinates, as integers: exact keys for the linking
    if side == _T:
        return (2 * r0, 2 * c0 + 1)
    if side == _B:
        return (2 * r0 + 2, 2 * c0 + 1)
    if side == _L:
        return (2 * r0 + 1, 2 * c0)
    return (2 * r0 + 1, 2 * c0 + 2)


def assemble(cells, cases, w):
    """Contours of ONE label from its cell records in raster order: list of (N, 2) float64 arrays of
    (row, col) points, in skimage's contour order."""
    current = 0
    contours, starts, ends = {}, {}, {}
    for cell, case in zip(cells.tolist(), cases.tolist()):
        r0, c0 = divmod(cell, w)
        for a, b in CASE_SEGMENTS[case]:
            p, q = _side_point(r0, c0, a), _side_point(r0, c0, b)
            tail, tail_num = starts.pop(q, (None, None))
            head, head_num = ends.pop(p, (None, None))
            if tail is not None and head is not None:
                if tail is head:
                    head.append(q)
                elif tail_num > head_num:
                    head.extend(tail)
                    contours.pop(tail_num, None)
                    starts[head[0]] = (head, head_num)
                    ends[head[-1]] = (head, head_num)
                else:
                    tail.extendleft(reversed(head))
                    starts.pop(head[0], None)
                    contours.pop(head_num, None)
                    starts[tail[0]] = (tail, tail_num)
                    ends[tail[-1]] = (tail, tail_num)
            elif tail is None and head is None:
                d = deque((p, q))
                contours[current] = d
                starts[p] = (d, current)
                ends[q] = (d, current)
                current += 1
            elif head is None:
                tail.appendleft(p)
                starts[p] = (tail, tail_num)
            else:
                head.append(q)
                ends[q] = (head, head_num)
    return [np.array(c, dtype=np.float64) / 2.0 for _, c in sorted(contours.items())]


def contours_of_crop(records, n_records, w):
    """{label: [contours]} of one crop from its record slice (uint32 [n][2])."""
    rec = np.asarray(records[:n_records]).reshape(-1, 2)
    if rec.shape[0] == 0:
        return {}
    cell, lab, case = rec[:, 0].astype(np.int64), (rec[:, 1] >> 4).astype(np.int64), (rec[:, 1] & 15).astype(np.int64)
    order = np.lexsort((cell, lab))                           # by label, then raster order of the cells
    cell, lab, case = cell[order], lab[order], case[order]
    cuts = np.flatnonzero(np.diff(lab)) + 1
    out = {}
    for lo, hi in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [lab.shape[0]]])):
        out[int(lab[lo])] = assemble(cell[lo:hi], case[lo:hi], w)
    return out
