"""Batched baselines: the one-launch MLP forward against the reference's own MTFNN / PPOAgent classes
evaluated on the bundled checkpoints (tests/golden/baselines.npz, made by oracle/make_golden_baselines.py)."""
import numpy as np
import pytest
import torch

from baseline_cases import CASES, build_case, golden

pytestmark = pytest.mark.gpu
TOL = 2e-6       # fp32 end to end; expf/tanhf vs torch's CPU routines


@pytest.mark.parametrize("name", sorted(CASES))
def test_baseline_forward_matches_reference(name):
    z = golden()
    model = build_case(name, z).cuda()
    x = torch.from_numpy(z[f"{name}.x"]).cuda()
    if name.startswith("ppo"):
        value, dist = model(x)
        assert np.abs(value.cpu().numpy() - z[f"{name}.value"]).max() < TOL * max(1.0, np.abs(z[f"{name}.value"]).max())
        assert np.abs(dist.mean.cpu().numpy() - z[f"{name}.mu"]).max() < TOL * max(1.0, np.abs(z[f"{name}.mu"]).max())
        np.testing.assert_allclose(dist.stddev.detach().cpu().numpy(), z[f"{name}.std"], rtol=1e-6)
    else:
        y = model(x).cpu().numpy()
        assert y.shape == z[f"{name}.y"].shape
        assert np.abs(y - z[f"{name}.y"]).max() < TOL


def test_baseline_forward_large_batch_is_row_independent():
    """65 537 rows (ragged against the block size): every row equals the same row evaluated in a batch of 300."""
    z = golden()
    model = build_case("mtfnn_nu", z).cuda()
    x = torch.from_numpy(z["mtfnn_nu.x"]).cuda()
    big = x.repeat(219, 1)[:65537]
    y = model(big)
    ref = model(x).repeat(219, 1)[:65537]
    assert torch.equal(y, ref)
    assert torch.allclose(y[:, 2:].sum(dim=1), torch.ones(65537, device="cuda"), atol=1e-5)


def test_baseline_forward_empty_batch_and_cpu_refusal():
    from diffsg_b200 import _lib
    z = golden()
    model = build_case("mtfnn_co", z).cuda()
    assert model(torch.zeros(0, 9, device="cuda")).shape == (0, 3)
    with pytest.raises(_lib.DiffsgError):
        model(torch.zeros(4, 9))
