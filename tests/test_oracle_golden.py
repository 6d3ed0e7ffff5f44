"""Pin the oracle: it must reproduce the reference's own outputs (tests/golden/*.npz).

The fixtures were minted by tests/golden/make_golden.py from the unmodified
reference on this container's CPU.  The oracle issues the same ATen ops, so on
the CPU the agreement is to the last few ulps; tolerances below leave room for
a different BLAS/threading on another host.
"""
import pytest
import torch

from helpers import Golden, golden_cases, max_abs, rel_l2
from oracle import loss_path as O

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 7


@pytest.mark.parametrize("case", CASES)
def test_masks_match_reference(case):
    g = Golden(case)
    fids = O.frame_ids_from_ordering(g.ordering)
    m = O.sub_batch_masks(g.ordering, fids, O.initial_valid_frames(g.ordering), g.trimin)
    assert m.valid_frames == g.valid_frames
    for attr, ref in g.masks.items():
        mine = getattr(m, attr)
        assert {str(k) for k in mine} == set(ref), (attr, sorted(map(str, mine)), sorted(ref))
        for k, v in mine.items():
            assert list(v) == ref[str(k)], (attr, k)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference(case):
    g = Golden(case)
    out, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    for k, v in g.losses.items():
        assert abs(float(out[k]) - v) <= 2e-7 * max(1.0, abs(v)), (k, float(out[k]), v)
    # warped images / depth planes of the first scale
    for k, ref in g.ref_out.items():
        assert max_abs(g.outputs[k], ref) <= 1e-6, k
    out["loss"].backward()
    for k, ref in g.grads.items():
        got = g.params[k].grad
        got = torch.zeros_like(ref) if got is None else got
        assert rel_l2(got, ref) <= 1e-6, (k, rel_l2(got, ref))


def test_argmin_shapes():
    g = Golden("trimin_mixed")
    out, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    n = sum(a.shape[0] for a in aux["argmin"][0])
    assert n == len(g.baselines)
