"""Host-side sub-batch bookkeeping: ``ordering`` -> integer index tables.

The reference regroups a mixed-baseline batch with Python lists of booleans
(``Trainer.valid_frames_trimin``, reference ``trainer.py:888-981``) and then
copies whole tensors with boolean indexing (``trainer.py:426-429,501-540``).
The fused kernels instead consume one small int32 table per batch that says,
for every target sample, which source rows it is warped from, with which
pose, in which candidate order, and which noise plane its identity terms
get.  Nothing here touches the GPU; tables are cached per ``ordering``.

Vocabulary (reference ``trainer.py:867-886``):
  * ``ordering[b]`` is ``[0, m, -m]`` (temporal baseline m>=1) or ``[0, 's']``.
  * frame *stack* ``("color", f, 0)`` holds only the rows of samples with
    baseline >= |f|; the ``'s'`` stack holds samples whose baseline is 's',1,2.
  * a *group* g collects the samples whose own baseline is g; its candidate
    list (order matters: ties go to the lowest index) follows
    ``trainer.py:549-555`` (plain) or ``Trainer.x_min_opt`` ``:983-1100``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np

Frame = Union[int, str]

MAX_REP = 12   # 6 sources + their 6 error-induced twins (x_min_opt, g>=3, decomp)
MAX_IDENT = 6
MAX_FRAMES = 16  # -7..7 plus 's'
MAX_GROUPS = 8   # baselines 1..7 plus 's'

# columns of the per-candidate table rows
REP_COLS = 4     # frame slot, stack row, pose row, K row
IDENT_COLS = 2   # frame slot, stack row
HDR_COLS = 4     # n_rep, n_ident, noise slot, noise row


def baseline_of(entry) -> Frame:
    """``[0, m, -m]`` -> m ; ``[0, 's']`` -> 's'."""
    return entry[1]


def frame_sort_key(f: Frame):
    return (1 << 30) if f == "s" else (2 * abs(f) + (1 if f < 0 else 0))


@dataclass
class LossPlan:
    """Everything the fused path needs to know about one batch layout."""

    batch: int
    trimin: bool
    decomp: bool
    baselines: List[Frame]
    frames: List[Frame]                       # slot order for image stacks
    groups: List[Frame]                       # noise / concat order
    sel: Dict[Frame, List[int]]               # frame -> target samples warped from it
    stack_row: Dict[Frame, List[int]]         # frame -> row in ("color", f, 0) per sel entry
    pose_off: Dict[Frame, int]                # frame -> first row in the packed P tensor
    pose_err_off: Dict[Frame, int]            # same for the error-induced poses (decomp)
    n_pose: int
    group_members: Dict[Frame, List[int]]
    rep: List[List[Tuple[Frame, bool]]]       # per sample: (frame, is_error_induced)
    ident: List[List[Frame]]                  # per sample
    hdr: np.ndarray = field(repr=False, default=None)      # (B, HDR_COLS) int32
    rep_tab: np.ndarray = field(repr=False, default=None)  # (B, MAX_REP, REP_COLS) int32
    ident_tab: np.ndarray = field(repr=False, default=None)  # (B, MAX_IDENT, IDENT_COLS) int32

    @property
    def warps_per_scale(self) -> int:
        """Sample-frames warped per scale = what ``px_pairs`` counts."""
        n = sum(len(v) for v in self.sel.values())
        if self.decomp:
            n += sum(len(v) for f, v in self.sel.items() if f != "s")
        return n

    @property
    def max_rep(self) -> int:
        return max(len(r) for r in self.rep)

    def frame_slot(self, f: Frame) -> int:
        return self.frames.index(f)

    def pose_slices(self):
        """[(frame, is_err, start, stop)] in packed-P order."""
        out = []
        for f in self.frames:
            out.append((f, False, self.pose_off[f], self.pose_off[f] + len(self.sel[f])))
        if self.decomp:
            for f in self.frames:
                if f != "s":
                    out.append((f, True, self.pose_err_off[f], self.pose_err_off[f] + len(self.sel[f])))
        return out


def _tri_sources(g: Frame) -> List[Frame]:
    """Source frames of group g under tri-minimisation (``x_min_opt``)."""
    if g == "s":
        return ["s"]
    if g == 1:
        return [1, -1, "s"]
    if g == 2:
        return [2, -2, 1, -1, "s"]
    return [g, -g, g - 1, -(g - 1), g - 2, -(g - 2)]


def _plain_sources(g: Frame) -> List[Frame]:
    return ["s"] if g == "s" else [g, -g]


def build_plan(ordering: Sequence[Sequence[Frame]], trimin: bool = False, decomp: bool = False,
               groups: Sequence[Frame] = None, s_stack_rows: int = None) -> LossPlan:
    """Derive the index tables for one batch.

    ``groups`` fixes the order in which per-group noise is drawn/concatenated
    (the reference iterates a ``set``, ``trainer.py:516-523``); default is the
    ascending temporal baselines followed by 's'.
    ``s_stack_rows``: number of rows actually present in ``("color","s",0)``;
    used to tell the two layouts apart in plain mode (rows only for 's'-baseline
    samples vs. rows for every sample with baseline <= 2).
    """
    B = len(ordering)
    base = [baseline_of(o) for o in ordering]
    present = sorted({g for g in base}, key=frame_sort_key)
    if groups is None:
        groups = present
    else:
        groups = [g for g in groups if g in present]
        assert sorted(groups, key=frame_sort_key) == present, (groups, present)
    decomp = bool(decomp and trimin)  # plain min ignores the error-induced planes (trainer.py:549)

    sources = _tri_sources if trimin else _plain_sources
    per_sample_src = [sources(g) for g in base]

    frames = sorted({f for srcs in per_sample_src for f in srcs}, key=frame_sort_key)
    assert len(frames) <= MAX_FRAMES and len(groups) <= MAX_GROUPS

    # rows of the compacted image stacks (custom_collate, trainer.py:867-886)
    def numeric(b):
        return base[b] != "s"

    s_members_tri = [b for b in range(B) if base[b] in ("s", 1, 2)]
    s_members_own = [b for b in range(B) if base[b] == "s"]
    if s_stack_rows is None:
        s_members = s_members_tri if trimin else s_members_own
    elif s_stack_rows == len(s_members_tri):
        s_members = s_members_tri
    elif s_stack_rows == len(s_members_own):
        s_members = s_members_own
    else:
        raise ValueError(f"'s' stack has {s_stack_rows} rows; expected "
                         f"{len(s_members_own)} or {len(s_members_tri)}")

    def stack_members(f):
        if f == "s":
            return s_members
        return [b for b in range(B) if numeric(b) and base[b] >= abs(f)]

    sel, stack_row = {}, {}
    for f in frames:
        members = stack_members(f)
        chosen = [b for b in range(B) if f in per_sample_src[b]]
        for b in chosen:
            if b not in members:
                raise ValueError(f"sample {b} (baseline {base[b]}) needs frame {f} "
                                 f"but has no row in its stack")
        sel[f] = chosen
        stack_row[f] = [members.index(b) for b in chosen]

    pose_off, off = {}, 0
    for f in frames:
        pose_off[f] = off
        off += len(sel[f])
    pose_err_off = {}
    if decomp:
        for f in frames:
            if f != "s":
                pose_err_off[f] = off
                off += len(sel[f])
    n_pose = off

    group_members = {g: [b for b in range(B) if base[b] == g] for g in groups}

    rep, ident = [], []
    for b in range(B):
        srcs = per_sample_src[b]
        r = [(f, False) for f in srcs]
        if decomp:
            r += [(f, True) for f in srcs if f != "s"]
        rep.append(r)
        ident.append(list(srcs))

    hdr = np.zeros((B, HDR_COLS), np.int32)
    rep_tab = np.zeros((B, MAX_REP, REP_COLS), np.int32)
    ident_tab = np.zeros((B, MAX_IDENT, IDENT_COLS), np.int32)
    for b in range(B):
        g = base[b]
        hdr[b] = (len(rep[b]), len(ident[b]), groups.index(g), group_members[g].index(b))
        for k, (f, is_err) in enumerate(rep[b]):
            i = sel[f].index(b)
            pose = (pose_err_off[f] if is_err else pose_off[f]) + i
            rep_tab[b, k] = (frames.index(f), stack_row[f][i], pose, i)
        for k, f in enumerate(ident[b]):
            i = sel[f].index(b)
            ident_tab[b, k] = (frames.index(f), stack_row[f][i])

    return LossPlan(batch=B, trimin=trimin, decomp=decomp, baselines=base, frames=frames,
                    groups=list(groups), sel=sel, stack_row=stack_row, pose_off=pose_off,
                    pose_err_off=pose_err_off, n_pose=n_pose, group_members=group_members,
                    rep=rep, ident=ident, hdr=hdr, rep_tab=rep_tab, ident_tab=ident_tab)


def shard_ordering(ordering: Sequence[Sequence[Frame]], rank: int, world: int):
    """Batch-shard a layout: rank r owns samples [r*B/world, (r+1)*B/world).

    Returns the local ordering and, per frame stack, the rows of the global
    compacted stack that belong to this rank (so a loader can slice them).
    """
    B = len(ordering)
    assert B % world == 0, "global batch must divide evenly across ranks"
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    local = [list(o) for o in ordering[lo:hi]]
    base = [baseline_of(o) for o in ordering]
    frames = sorted({f for o in ordering for f in o if f != 0}, key=frame_sort_key)
    max_m = max([m for m in base if m != "s"], default=0)
    all_frames = [f for f in range(-max_m, max_m + 1) if f != 0]
    if any(m in ("s", 1, 2) for m in base):
        all_frames.append("s")
    rows = {}
    for f in all_frames:
        if f == "s":
            members = [b for b in range(B) if base[b] in ("s", 1, 2)]
        else:
            members = [b for b in range(B) if base[b] != "s" and base[b] >= abs(f)]
        rows[f] = [i for i, b in enumerate(members) if lo <= b < hi]
    del frames
    return local, rows
