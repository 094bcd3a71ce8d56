"""CPU: the oracle restatement (oracle/dpt_oracle.py) against the golden vectors produced by the real reference
(oracle/make_golden.py, run in the build container). The reference ships no tests of its own, so
these fixtures are what pins the oracle."""
import os

import pytest
import torch

from oracle import dpt_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    fix = torch.load(os.path.join(GOLDEN, name))
    sd = fix.get("state_dict") or O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    return fix, sd


@pytest.mark.parametrize("name", ["tiny_a.pt", "tiny_b.pt"])
def test_oracle_matches_reference_every_stage(name):
    fix, sd = _fixture(name)
    st = O.forward(sd, fix["img"], return_stages=True)
    assert tuple(st["grid_hw"]) == tuple(fix["grid_hw"])
    torch.testing.assert_close(st["tokens"], fix["tokens"], rtol=0, atol=1e-6)
    for a, b in zip(st["taps"], fix["taps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-5)
    for a, b in zip(st["maps"], fix["maps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=5e-5)
    torch.testing.assert_close(st["fused"], fix["fused"], rtol=0, atol=1e-4)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("name", ["vits_a.pt", "vits_b.pt"])
def test_oracle_matches_reference_vits(name):
    from oracle.make_golden import state_dict_checksum, sub

    fix, sd = _fixture(name)
    assert state_dict_checksum(sd) == fix["sd_checksum"]
    st = O.forward(sd, fix["img"], return_stages=True)
    torch.testing.assert_close(sub(st["tokens"]), fix["tokens_sub"], rtol=0, atol=1e-5)
    for a, b in zip(st["taps"], fix["taps_sub"]):
        torch.testing.assert_close(sub(a), b, rtol=0, atol=1e-4)
    for a, b in zip(st["maps"], fix["maps_sub"]):
        torch.testing.assert_close(sub(a), b, rtol=0, atol=2e-4)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=5e-4)


def test_oracle_manual_attention_equals_sdpa():
    fix, sd = _fixture("tiny_a.pt")
    a = O.forward(sd, fix["img"], use_sdpa=True)
    b = O.forward(sd, fix["img"], use_sdpa=False)
    torch.testing.assert_close(a, b, rtol=0, atol=5e-5)


def test_oracle_config_inference():
    sd = O.make_synthetic_state_dict("vits", seed=11)
    cfg = O.infer_config(sd)
    assert cfg["features_per_token"] == 384 and cfg["num_blocks"] == 12 and cfg["num_heads"] == 6
    assert cfg["reassembly_features_list"] == [48, 96, 192, 384] and cfg["fusion_channels"] == 64
    assert cfg["base_patch_grid_hw"] == (37, 37) and cfg["patch_size_px"] == 14


def test_oracle_rejects_odd_grid_like_the_reference():
    # 518 / 14 = 37 (odd): the reference raises inside fusion; so does the restatement
    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    with pytest.raises(RuntimeError):
        O.forward(sd, O.make_input(1, 42, 42))


@pytest.mark.parametrize("name", ["beit_tiny_a.pt", "beit_tiny_b.pt"])
def test_beit_oracle_matches_reference_every_stage(name):
    from oracle.make_golden import state_dict_checksum

    fix = torch.load(os.path.join(GOLDEN, name))
    sd = O.make_synthetic_state_dict_beit(fix["sd_name"], fix["sd_seed"])
    assert state_dict_checksum(sd) == fix["sd_checksum"]
    st = O.forward_beit(sd, fix["img"], return_stages=True)
    assert tuple(st["grid_hw"]) == tuple(fix["grid_hw"])
    torch.testing.assert_close(st["tokens"], fix["tokens"], rtol=0, atol=1e-6)
    for a, b in zip(st["taps"], fix["taps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=5e-5)
    for a, b in zip(st["maps"], fix["maps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-4)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=2e-4)


def test_beit_relative_position_index_special_rows():
    idx = O.beit_relative_position_index((3, 4))
    n = 13
    top = (2 * 3 - 1) * (2 * 4 - 1) - 1
    assert idx.shape == (n, n)
    assert (idx[0, 1:] == top + 1).all() and (idx[1:, 0] == top + 2).all() and idx[0, 0] == top + 3
    assert idx[1:, 1:].min() == 0 and idx[1:, 1:].max() == top
    assert (idx[1:, 1:].diagonal() == (3 - 1) * (2 * 4 - 1) + (4 - 1)).all()  # zero offset


@pytest.mark.parametrize("name", ["swinv2_micro_a.pt", "swinv2_micro_b.pt"])
def test_swinv2_oracle_matches_reference_every_stage(name):
    from oracle.make_golden import state_dict_checksum

    fix = torch.load(os.path.join(GOLDEN, name))
    sd = O.make_synthetic_state_dict_swinv2(fix["sd_name"], fix["sd_seed"])
    assert state_dict_checksum(sd) == fix["sd_checksum"]
    st = O.forward_swinv2(sd, fix["img"], return_stages=True)
    assert tuple(st["grid_hw"]) == tuple(fix["grid_hw"])
    torch.testing.assert_close(st["tokens"], fix["tokens"], rtol=0, atol=1e-5)
    for a, b in zip(st["taps"], fix["taps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-4)
    for a, b in zip(st["maps"], fix["maps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-4)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=5e-4)


def test_swin_window_and_shift_rules():
    # adjust_window_and_shift_sizes (windowed_attention.py:345-388): SwinV2-L @384 (grids 96/48/24/12, target 24)
    assert O.swin_window_and_shift((96, 96), (24, 24)) == ((24, 24), (12, 12))
    assert O.swin_window_and_shift((24, 24), (24, 24)) == ((24, 24), (0, 0))
    assert O.swin_window_and_shift((12, 12), (24, 24)) == ((12, 12), (0, 0))
    # non-tiling grid: closest divisor of the grid in [win/2, 2 win)
    assert O.swin_window_and_shift((32, 20), (8, 8)) == ((8, 10), (4, 5))
    m = O.swin_shift_mask((16, 16), (8, 8), (4, 4))
    assert m.shape == (4, 64, 64) and set(m.unique().tolist()) == {-100.0, 0.0}
    assert (m[0] == 0).all()  # the top-left window never wraps
    assert O.swin_shift_mask((8, 8), (8, 8), (0, 0)) is None
