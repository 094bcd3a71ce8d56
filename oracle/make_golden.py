#!/usr/bin/env python3
"""
Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference, build container only)
through its own loader (make_dpt_from_state_dict on a synthetic upstream-format checkpoint) and its
own per-stage calls (simple_examples/internal_features.py:38-44). Run:  python oracle/make_golden.py

Fixtures (all fp32, CPU):
  tiny_*.pt   - a 4-block F=128 model: full checkpoint + input + every stage tensor (small enough to commit whole)
  vits_*.pt   - the real ViT-S architecture at 112x112 / 140x84: checkpoint is regenerated from its seed at test time
                (oracle.make_synthetic_state_dict) and guarded by a checksum; stage tensors are stored sub-sampled,
                the depth map whole.
"""
import hashlib
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dpt_oracle as O  # noqa: E402


def state_dict_checksum(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def run_reference(sd: dict, img: torch.Tensor, enable_optimizations: bool):
    from muggled_dpt.make_dpt import make_dpt_from_state_dict

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "depth_anything_v2_synthetic.pth")  # 'v2' in the name -> depthanythingv2 (make_dpt.py:98)
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path, enable_cache=False, enable_optimizations=enable_optimizations)
    model = model.float().eval()
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        taps = model.imgencoder(tokens, grid_hw)
        maps = model.reassemble(*taps, grid_hw)
        fused = model.fusion(*maps)
        depth = model.head(fused)
        depth_whole = model(img)
    assert torch.equal(depth, depth_whole)
    return cfg, {"tokens": tokens, "taps": taps, "maps": maps, "fused": fused, "depth": depth, "grid_hw": tuple(grid_hw)}


def sub(t: torch.Tensor, n: int = 4096) -> torch.Tensor:
    """deterministic sub-sample of a tensor (flat stride), to keep fixtures small"""
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].clone()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)

    # ---- tiny model: everything stored
    for tag, (H, W) in {"a": (56, 56), "b": (84, 112)}.items():
        sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
        img = O.make_input(2, H, W, seed=7)
        cfg, ref = run_reference(sd, img, enable_optimizations=True)
        _, ref_manual = run_reference(sd, img, enable_optimizations=False)
        print(f"tiny_{tag}: sdpa-vs-manual depth max abs diff", (ref["depth"] - ref_manual["depth"]).abs().max().item())
        fix = {
            "config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()},
            "state_dict": sd if tag == "a" else None,
            "sd_seed": 3, "sd_name": "tiny", "sd_base_grid": 5,
            "sd_checksum": state_dict_checksum(sd),
            "img": img,
            "tokens": ref["tokens"], "taps": list(ref["taps"]), "maps": list(ref["maps"]),
            "fused": ref["fused"], "depth": ref["depth"], "grid_hw": ref["grid_hw"],
        }
        torch.save(fix, os.path.join(out_dir, f"tiny_{tag}.pt"))
        for k in ("tokens", "fused", "depth"):
            print(f"  {k}: std {ref[k].std().item():.4f} mean {ref[k].mean().item():.4f}")

    # ---- real ViT-S architecture, small images; checkpoint regenerated from seed
    for tag, (B, H, W) in {"a": (1, 112, 112), "b": (2, 140, 84)}.items():
        sd = O.make_synthetic_state_dict("vits", seed=11)
        img = O.make_input(B, H, W, seed=5)
        cfg, ref = run_reference(sd, img, enable_optimizations=True)
        fix = {
            "config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()},
            "sd_seed": 11, "sd_name": "vits", "sd_base_grid": 37,
            "sd_checksum": state_dict_checksum(sd),
            "img": img,
            "tokens_sub": sub(ref["tokens"]), "taps_sub": [sub(t) for t in ref["taps"]],
            "maps_sub": [sub(t) for t in ref["maps"]], "fused_sub": sub(ref["fused"]),
            "depth": ref["depth"], "grid_hw": ref["grid_hw"],
        }
        torch.save(fix, os.path.join(out_dir, f"vits_{tag}.pt"))
        print(f"vits_{tag}: depth std {ref['depth'].std().item():.4f} mean {ref['depth'].mean().item():.4f} "
              f"min {ref['depth'].min().item():.4f}; taps std {[round(t.std().item(), 3) for t in ref['taps']]}")

    # ---- MiDaS v3.1 BEiT (tiny synthetic config; everything stored except the checkpoint, regenerated from its seed)
    for tag, (B, H, W) in {"a": (2, 96, 96), "b": (1, 96, 128)}.items():  # a: base grid (LUT resize = identity)
        sd = O.make_synthetic_state_dict_beit("beit_tiny", seed=5)
        img = O.make_input(B, H, W, seed=3)
        cfg, ref = run_reference(sd, img, enable_optimizations=True)
        fix = {
            "model_type": "beit",
            "config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()},
            "sd_seed": 5, "sd_name": "beit_tiny", "sd_checksum": state_dict_checksum(sd),
            "img": img,
            "tokens": ref["tokens"], "taps": list(ref["taps"]), "maps": list(ref["maps"]),
            "fused": ref["fused"], "depth": ref["depth"], "grid_hw": ref["grid_hw"],
        }
        torch.save(fix, os.path.join(out_dir, f"beit_tiny_{tag}.pt"))
        print(f"beit_tiny_{tag}: depth std {ref['depth'].std().item():.4f} mean {ref['depth'].mean().item():.4f}; "
              f"taps std {[round(t.std().item(), 3) for t in ref['taps']]}")

    # ---- MiDaS v3.1 SwinV2 (micro synthetic config: window 8, 4 stages x 2 blocks; b is non-square -> 8x10 windows)
    for tag, (B, H, W) in {"a": (2, 128, 128), "b": (1, 128, 160)}.items():
        sd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=4)
        img = O.make_input(B, H, W, seed=3)
        cfg, ref = run_reference(sd, img, enable_optimizations=True)
        fix = {
            "model_type": "swinv2",
            "config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()},
            "sd_seed": 4, "sd_name": "swinv2_micro", "sd_checksum": state_dict_checksum(sd),
            "img": img,
            "tokens": ref["tokens"], "taps": list(ref["taps"]), "maps": list(ref["maps"]),
            "fused": ref["fused"], "depth": ref["depth"], "grid_hw": ref["grid_hw"],
        }
        torch.save(fix, os.path.join(out_dir, f"swinv2_micro_{tag}.pt"))
        print(f"swinv2_micro_{tag}: depth std {ref['depth'].std().item():.4f} mean {ref['depth'].mean().item():.4f}; "
              f"taps std {[round(t.std().item(), 3) for t in ref['taps']]}")

    # ---- oracle vs reference, right here
    for name in sorted(os.listdir(out_dir)):
        if not name.endswith(".pt"):
            continue
        fix = torch.load(os.path.join(out_dir, name))
        if fix.get("model_type") == "beit":
            sd = O.make_synthetic_state_dict_beit(fix["sd_name"], fix["sd_seed"])
            st = O.forward_beit(sd, fix["img"], return_stages=True)
        elif fix.get("model_type") == "swinv2":
            sd = O.make_synthetic_state_dict_swinv2(fix["sd_name"], fix["sd_seed"])
            st = O.forward_swinv2(sd, fix["img"], return_stages=True)
        else:
            sd = fix.get("state_dict") or O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
            st = O.forward(sd, fix["img"], return_stages=True)
        print(name, "oracle-vs-reference depth max abs diff:", (st["depth"] - fix["depth"]).abs().max().item())


if __name__ == "__main__":
    main()
