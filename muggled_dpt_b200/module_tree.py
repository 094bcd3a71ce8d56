"""
Module tree for the reference's `experiments/`: hook points and per-block callables.

The forward path is a launch plan inside libdpt_b200.so, not a tree of nn.Modules. The reference's analysis scripts
however reach into the tree:

  * `ModelOutputCapture(model, torch.nn.Softmax)` registers forward hooks on every nn.Softmax and expects the attention
    probabilities of each block (demo_helpers/model_capture.py:15-61, experiments/attention_visualization.py:324-332);
  * `ModelOutputCapture(model, TransformerBlock)` does the same for whole encoder blocks
    (experiments/block_norm_visualization.py:265-300);
  * `model.fusion.blocks[i](reassembly_map, previous_fusion)` runs one fusion block (experiments/fusion_scaling.py:330-333).

So the encoder and fusion stages are nn.Modules with the same shape: `imgencoder.blocks[i]` (TransformerBlock /
SwinTransformerBlock) with `.attn.softmax` (an nn.Softmax subclass), and `fusion.blocks[i]` (FusionBlock, callable).
When a hook is registered on any of the encoder's hook points, `imgencoder(tokens, grid)` goes through
`dpt_encoder_capture`, which runs the same kernels plus a debug kernel that materialises the probabilities, and then
calls the hooked modules with the captured tensors so the hooks fire exactly as they do in the reference. Without hooks
nothing changes (same single C call as before).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _native as N


class AttentionSoftmax(torch.nn.Softmax):
    """hook point: receives the block's attention probabilities [B, heads, N, N] (SwinV2: [B*windows, heads, A, A]),
    already normalised by the device kernel - transformer_block.py:101,132"""

    def __init__(self):
        super().__init__(dim=-1)

    def forward(self, attention_probabilities: torch.Tensor) -> torch.Tensor:
        return attention_probabilities


class _AttentionHookPoint(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.softmax = AttentionSoftmax()


class TransformerBlock(torch.nn.Module):
    """hook point for one encoder block (transformer_block.py:21-65; v31_beit/image_encoder_model.py:233-251). The block
    itself runs inside the encoder's launch plan; calling the module fires its hooks with the block's output tokens."""

    def __init__(self, index: int):
        super().__init__()
        self.index = index
        self.attn = _AttentionHookPoint()

    def forward(self, block_output_tokens: torch.Tensor) -> torch.Tensor:
        return block_output_tokens


class SwinTransformerBlock(TransformerBlock):
    """v31_swinv2/image_encoder_model.py:164-225"""


def _has_hooks(m: torch.nn.Module) -> bool:
    return len(m._forward_hooks) > 0 or len(m._forward_pre_hooks) > 0


def swin_window_and_shift(patch: int, target: int) -> tuple[int, int]:
    """adjust_window_and_shift_sizes, one axis - v31_swinv2/components/windowed_attention.py:345-388"""
    win = min(target, patch)
    if patch % win != 0:
        cands = [d for d in range(max(1, win // 2), 2 * win) if patch % d == 0]
        win = min(cands, key=lambda d: abs(patch - d))
    return win, (0 if patch <= win else win // 2)


class ImageEncoder(torch.nn.Module):
    """`model.imgencoder(tokens, grid_hw) -> 4 taps` (image_encoder_model.py:80-94) + the hook points above"""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "_model", model)  # not a sub-module: avoids a reference cycle in model.modules()
        cfg = model.config
        if model.model_type == "swinv2":
            n_blocks, cls = sum(cfg["layers_per_stage"]), SwinTransformerBlock
        else:
            n_blocks, cls = cfg["num_blocks"], TransformerBlock
        self.blocks = torch.nn.ModuleList(cls(i) for i in range(n_blocks))

    def _block_shapes(self, B, gh, gw):
        """per block: (probabilities shape, output tokens shape)"""
        m = self._model
        cfg = m.config
        if m.model_type != "swinv2":
            n, f, h = gh * gw + 1, cfg["features_per_token"], cfg["num_heads"]
            return [((B, h, n, n), (B, n, f))] * cfg["num_blocks"]
        out = []
        for st in range(4):
            sgh, sgw, f, h = gh >> st, gw >> st, cfg["features_per_stage"][st], cfg["heads_per_stage"][st]
            wh, _ = swin_window_and_shift(sgh, cfg["window_size_hw"][0])
            ww, _ = swin_window_and_shift(sgw, cfg["window_size_hw"][1])
            a, nw = wh * ww, (sgh // wh) * (sgw // ww)
            out += [((B * nw, h, a, a), (B, sgh * sgw, f))] * cfg["layers_per_stage"][st]
        return out

    def forward(self, patch_tokens: torch.Tensor, patch_grid_hw):
        m = self._model
        want_probs = [_has_hooks(b.attn.softmax) for b in self.blocks]
        want_out = [_has_hooks(b) for b in self.blocks]
        if not any(want_probs) and not any(want_out):
            return m._stage_encoder(patch_tokens, patch_grid_hw)
        gh, gw = int(patch_grid_hw[0]), int(patch_grid_hw[1])
        B = patch_tokens.shape[0]
        device, dtype = m._require_ready()
        shapes = self._block_shapes(B, gh, gw)
        probs = [torch.empty(s[0], dtype=dtype, device=device) if w else None for s, w in zip(shapes, want_probs)]
        outs = [torch.empty(s[1], dtype=dtype, device=device) if w else None for s, w in zip(shapes, want_out)]
        taps = m._stage_encoder(patch_tokens, patch_grid_hw, capture=(probs, outs))
        for blk, p, o in zip(self.blocks, probs, outs):  # fire the hooks in execution order, like the reference
            if p is not None:
                blk.attn.softmax(p)
            if o is not None:
                blk(o)
        return taps


class FusionBlock(torch.nn.Module):
    """`model.fusion.blocks[i]`: one fusion block as a callable (fusion_model.py:119-154; index 3 is the top-most block,
    called with the coarsest reassembly map only :89-114). BxCxHxW in / out like the reference (channels-last memory)."""

    def __init__(self, model, level: int):
        super().__init__()
        object.__setattr__(self, "_model", model)
        self.level = level

    def forward(self, reassembly_feature_map: torch.Tensor, previous_fusion_feature_map: torch.Tensor | None = None):
        m = self._model
        device, dtype = m._require_ready()
        if self.level < 3 and previous_fusion_feature_map is None:
            raise TypeError("FusionBlock.forward() missing the previous fusion feature map (only blocks[3] takes one input)")
        r = m._as_nhwc(reassembly_feature_map.to(dtype))
        B, Cc, h, w = r.shape
        prev = None
        if self.level < 3:
            prev = m._as_nhwc(previous_fusion_feature_map.to(dtype))
            if tuple(prev.shape) != (B, Cc, h, w):
                raise RuntimeError(f"fusion block {self.level}: size mismatch between the reassembly map {tuple(r.shape)} "
                                   f"and the previous fusion map {tuple(prev.shape)}")
        out = m._nhwc_empty(B, Cc, 2 * h, 2 * w)
        # any workspace sized for an image whose finest map is at least this large will do
        k0 = 1 if m.model_type == "swinv2" else 4
        p = m.config["patch_size_px"]
        mult = 8 if m.model_type == "swinv2" else 2
        gh = -(-(h << self.level) // k0)
        gw = -(-(w << self.level) // k0)
        gh, gw = -(-gh // mult) * mult, -(-gw // mult) * mult
        ws = m._get_workspace(B, gh * p, gw * p)
        with torch.cuda.device(device):
            rc = N.lib().dpt_fusion_block(m._handle, self.level, C.c_void_p(r.data_ptr()),
                                          C.c_void_p(prev.data_ptr()) if prev is not None else None,
                                          C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), B, h, w, m._stream())
        N.check(rc, m._handle, "dpt_fusion_block")
        return out


class FusionModel(torch.nn.Module):
    """`model.fusion(r1, r2, r3, r4) -> fused map` (fusion_model.py:55-80) + `.blocks[i]`"""

    def __init__(self, model):
        super().__init__()
        object.__setattr__(self, "_model", model)
        self.blocks = torch.nn.ModuleList(FusionBlock(model, lvl) for lvl in range(4))

    def forward(self, r1, r2, r3, r4):
        return self._model._stage_fusion(r1, r2, r3, r4)
