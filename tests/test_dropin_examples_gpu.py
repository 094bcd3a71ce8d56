"""-m gpu: the reference's own example scripts (simple_examples/depth_prediction.py and internal_features.py) run
against the product with ONLY the import swapped (`muggled_dpt.make_dpt` -> `muggled_dpt_b200.make_dpt`) and the two
path constants filled in. The scripts are the unmodified files oracle/build_ref.py archived under oracle/_ref (the archive
travels to the GPU box with the snapshot); nothing here reads /root/reference."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_example(script_name, sd, ckpt_name, image_hw):
    import cv2

    from oracle.build_ref import read_example, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref is absent (run oracle/build_ref.py where /root/reference exists)")
    text = read_example(script_name)
    assert text.count("from muggled_dpt.make_dpt import make_dpt_from_state_dict") == 1
    with tempfile.TemporaryDirectory() as td:
        rng = np.random.default_rng(3)
        img = rng.integers(0, 255, size=(image_hw[0], image_hw[1], 3), dtype=np.uint8)
        img_path, model_path = os.path.join(td, "frame.png"), os.path.join(td, ckpt_name)
        cv2.imwrite(img_path, img)
        torch.save(sd, model_path)
        text = text.replace("from muggled_dpt.make_dpt import make_dpt_from_state_dict",
                            "from muggled_dpt_b200.make_dpt import make_dpt_from_state_dict")
        text = text.replace('image_path = "/path/to/image.jpg"', f"image_path = {img_path!r}")
        text = text.replace('model_path = "/path/to/model.pth"', f"model_path = {model_path!r}")
        # the example's own "find the muggled_dpt folder" hack wants an importable package of that name: an empty stub
        os.makedirs(os.path.join(td, "stub", "muggled_dpt"))
        open(os.path.join(td, "stub", "muggled_dpt", "__init__.py"), "w").close()
        script = os.path.join(td, script_name)
        open(script, "w").write(text)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(td, "stub")]))
        r = subprocess.run([sys.executable, script], cwd=td, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    return r.stdout


def test_depth_prediction_example_runs_with_the_import_swapped():
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    out = _run_example("depth_prediction.py", sd, "depth_anything_v2_vits_synth.pth", (300, 400))
    print(out)
    # use_square_sizing=False, default max side 518 -> 392 x 504 (multiples of 2 * 14), depth [1, H, W]
    assert "Result shape: (1, 392, 504)" in out
    assert re.search(r"Result min: \d", out) and re.search(r"Result max: \d", out)
    for key in ("features_per_token: 384", "num_blocks: 12", "patch_size_px: 14", "base_patch_grid_hw: (37, 37)"):
        assert key in out, key


def test_internal_features_example_runs_with_the_import_swapped():
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    out = _run_example("internal_features.py", sd, "depth_anything_v2_vits_synth.pth", (300, 400))
    print(out)
    assert "Pre-encoded image shape: (1, 3, 392, 504)" in out
    assert "Patch grid height & width (28, 36)" in out
    assert "Patch embedding shape: (1, 1008, 384)" in out
    assert "Image encoding stage 1 shape: (1, 1009, 384)" in out
    assert "Reassembly 1 result shape: (1, 64, 112, 144)" in out
    assert "Reassembly 4 result shape: (1, 64, 14, 18)" in out
    assert "Fused feature map shape: (1, 64, 224, 288)" in out
    assert "Final output shape: (1, 392, 504)" in out
