"""
DPTModel - the reference's model wrapper surface (muggled_dpt/dpt_model.py:21-166) over libdpt_b200.so.

Same calls as the reference: model(image_bchw), .inference(bgr), .prepare_image_bgr(...), .verify_input(...),
.to(device=, dtype=, memory_format=), and the five stage attributes patch_embed / imgencoder / reassemble / fusion /
head with the tensor signatures of simple_examples/internal_features.py:38-44. Everything below this surface runs in
hand-written sm_100a kernels through the C ABI; PyTorch only owns the buffers. There is no CPU path: calling the model
before `.to("cuda")`, or asking for fp32, raises.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .module_tree import FusionModel, ImageEncoder
from .weights import pack_beit, pack_depthanything_v2, pack_swinv2

_TORCH_TO_DPT = {torch.float16: N.DPT_F16, torch.bfloat16: N.DPT_BF16}


class _Stage:
    """callable stage wrapper (keeps `model.patch_embed(x)` etc. working)"""

    def __init__(self, model: "DPTModel", fn):
        self._model = model
        self._fn = fn

    def __call__(self, *args, **kwargs):
        return self._fn(*args, **kwargs)


class _PatchEmbedStage(_Stage):
    # input normalisation constants - v2_depthanything/patch_embed.py:38-39 ; v31_beit/patch_embed.py:38-39
    NORMALISATION = {
        "depthanythingv2": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
        "depthanythingv1": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),  # v1_depthanything/patch_embed.py:38-39
        "beit": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
        "swinv2": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),  # v31_swinv2/patch_embed.py:39-40
    }

    @property
    def rgb_offset(self):
        return self.NORMALISATION[self._model.model_type][0]

    @property
    def rgb_stdev(self):
        return self.NORMALISATION[self._model.model_type][1]

    def prepare_image(self, image_bgr, max_side_length=None, use_square_sizing=True, interpolation_mode="bilinear"):
        """PatchEmbed.prepare_image - v2_depthanything/patch_embed.py:103-145 (host-side glue, outside forward())"""
        m = self._model
        patch = m.config["patch_size_px"]
        default_size = m.config["base_patch_grid_hw"][0] * patch
        tiling = round((8 if m.model_type == "swinv2" else 2) * patch)  # v31_swinv2/patch_embed.py:68
        if max_side_length is None:
            max_side_length = default_size
        img_h, img_w = image_bgr.shape[0:2]
        largest = max(img_h, img_w)
        scale = max_side_length / largest
        targ_hw = (largest, largest) if use_square_sizing else (img_h, img_w)
        scaled_hw = [max(1, round(side * scale / tiling)) * tiling for side in targ_hw]
        device, dtype = m._require_ready()
        if interpolation_mode != "bilinear":
            raise NotImplementedError("prepare_image: the device kernel implements the reference's default, bilinear")
        bgr = np.ascontiguousarray(image_bgr)
        if bgr.dtype != np.uint8 or bgr.ndim != 3 or bgr.shape[2] != 3:
            raise ValueError("prepare_image: expected an HxWx3 uint8 BGR image (cv2.imread format)")
        # one kernel: BGR->RGB, antialiased bilinear resize, /255, mean / std, NCHW 16-bit (csrc prepare_image_kernel)
        with torch.cuda.device(device):
            raw = torch.from_numpy(bgr).to(device, non_blocking=True)
            out = torch.empty((1, 3, scaled_hw[0], scaled_hw[1]), device=device, dtype=dtype)
            mean = (C.c_float * 3)(*self.rgb_offset)
            inv_std = (C.c_float * 3)(*[1.0 / v for v in self.rgb_stdev])
            N.check(
                N.lib().dpt_prepare_image(C.c_void_p(raw.data_ptr()), img_h, img_w, C.c_void_p(out.data_ptr()), scaled_hw[0],
                                          scaled_hw[1], C.byref(mean), C.byref(inv_std), _TORCH_TO_DPT[dtype], m._stream()),
                None, "dpt_prepare_image",
            )
            raw.record_stream(torch.cuda.current_stream(device))
        return out

    def verify_input(self, image_tensor_bchw) -> bool:
        """PatchEmbed.verify_input - v2_depthanything/patch_embed.py:149-165 (+ the even-grid rule the reference only
        hits as a RuntimeError inside fusion)"""
        b, c, h, w = image_tensor_bchw.shape
        p = self._model.config["patch_size_px"]
        assert c == 3, f"Bad channel count! Expected 3 got {c}"
        assert h % p == 0, f"Bad height! Image must have height ({h}) divisble by {p}"
        assert w % p == 0, f"Bad width! Image must have width ({w}) divisble by {p}"
        assert (h // p) % 2 == 0 and (w // p) % 2 == 0, (
            f"Bad size! The patch grid ({h // p}x{w // p}) must be even in both directions "
            f"(use multiples of {2 * p} px, e.g. 504 or 532 instead of 518)"
        )
        return True


class DPTModel(torch.nn.Module):
    def __init__(self, config: dict, state_dict: dict, strict_load: bool = True, model_type: str = "depthanythingv2"):
        super().__init__()
        self.config = dict(config)
        self.model_type = model_type
        if model_type == "swinv2":
            fs, hs = self.config["features_per_stage"], self.config["heads_per_stage"]
            if any(f != 32 * h for f, h in zip(fs, hs)) or any(b != a * 2 for a, b in zip(fs, fs[1:])):
                raise NotImplementedError("SwinV2 stages must have 32 features per head and double in width")
            self.config.setdefault("features_per_token", fs[0])
            self._packed_cpu = pack_swinv2(state_dict, self.config, strict=strict_load)
        elif self.config["features_per_token"] != 64 * self.config["num_heads"]:
            raise NotImplementedError("the attention kernel is built for 64 features per head")
        # fp32 CPU copies in kernel layouts; moved/cast by .to()
        elif model_type == "beit":
            self._packed_cpu = pack_beit(state_dict, self.config, strict=strict_load)
        else:
            self._packed_cpu = pack_depthanything_v2(state_dict, self.config, strict=strict_load)
        self._dev_weights: dict[str, torch.Tensor] = {}
        self._handle = None
        self._device = None
        self._dtype = torch.bfloat16
        self._workspace = None
        self._ws_key = None
        self._io = {}
        self.patch_embed = _PatchEmbedStage(self, self._stage_patch_embed)
        self.imgencoder = ImageEncoder(self)  # nn.Modules with hook points / per-block callables (module_tree.py)
        self.reassemble = _Stage(self, self._stage_reassemble)
        self.fusion = FusionModel(self)
        self.head = _Stage(self, self._stage_head)
        self.eval()

    # ------------------------------------------------------------------------------------------------- placement

    def to(self, *args, **kwargs):
        device = kwargs.get("device", None)
        dtype = kwargs.get("dtype", None)
        for a in args:
            if isinstance(a, bool):  # positional non_blocking (bool is an int: not a device index)
                continue
            if isinstance(a, torch.dtype):
                dtype = a
            elif isinstance(a, (str, torch.device, int)):
                device = a
        if dtype is not None:
            if dtype not in _TORCH_TO_DPT:
                raise RuntimeError(
                    f"muggled_dpt_b200 computes in bf16 or fp16 (fp32 accumulate); {dtype} is only available from the "
                    "reference / oracle CPU path"
                )
            self._dtype = dtype
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise RuntimeError("muggled_dpt_b200 has no CPU fallback: move the model to a B200 with .to('cuda')")
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            self._device = device
        if self._device is not None:
            self._materialise()
        return self

    def cuda(self, device=None):
        return self.to(device=torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def half(self):
        return self.to(dtype=torch.float16)

    def bfloat16(self):
        return self.to(dtype=torch.bfloat16)

    def float(self):
        return self.to(dtype=torch.float32)

    def parameters(self, recurse: bool = True):
        # the reference's verify_input peeks at next(model.parameters()) for device / dtype
        self._require_ready()
        yield self._dev_weights["patch.w"]

    def _require_ready(self):
        if self._handle is None:
            # The reference's examples use the model right after make_dpt_from_state_dict() (on the CPU, in fp32 -
            # simple_examples/depth_prediction.py:32-33). There is no CPU path here: with a GPU present the model places
            # itself on the current CUDA device in bf16 on first use, without one it raises.
            if not torch.cuda.is_available():
                raise RuntimeError("model is not on a GPU yet: call .to('cuda') first (there is no CPU fallback)")
            self.to(device="cuda")
        return self._device, self._dtype

    def _materialise(self):
        L = N.lib()
        self._release()
        with torch.cuda.device(self._device):
            cfg = N.DptConfig()
            cfg.variant = {"beit": N.VARIANT_BEIT, "swinv2": N.VARIANT_SWINV2}.get(self.model_type, N.VARIANT_DINOV2)
            cfg.dtype = _TORCH_TO_DPT[self._dtype]
            cfg.features_per_token = self.config["features_per_token"]
            cfg.fusion_channels = self.config["fusion_channels"]
            cfg.patch_size_px = self.config["patch_size_px"]
            cfg.base_grid_h, cfg.base_grid_w = self.config["base_patch_grid_hw"]
            cfg.is_metric = int(bool(self.config.get("is_metric", False)))
            cfg.taps_last4 = int(self.model_type == "depthanythingv1")
            cfg.mlp_swiglu = int(bool(self.config.get("is_giant", False)))
            if self.model_type == "swinv2":
                cfg.ln_eps = 1e-5  # torch default, all SwinV2 LayerNorms
                cfg.window_h, cfg.window_w = self.config["window_size_hw"]
                for i in range(4):
                    cfg.heads_per_stage[i] = self.config["heads_per_stage"][i]
                    cfg.layers_per_stage[i] = self.config["layers_per_stage"][i]
                    cfg.pretrained_window[i] = self.config["pretrained_window_sizes_per_stage"][i] or 0
                    cfg.reassembly_features[i] = self.config["features_per_stage"][i]
            else:
                cfg.ln_eps = 1e-6
                cfg.num_heads = self.config["num_heads"]
                cfg.num_blocks = self.config["num_blocks"]
                for i, r in enumerate(self.config["reassembly_features_list"]):
                    cfg.reassembly_features[i] = r
            handle = C.c_void_p()
            N.check(L.dpt_create(C.byref(cfg), C.byref(handle)), None, "dpt_create")
            self._handle = handle
            self._dev_weights = {}
            def set_weight(name, d, code):
                self._dev_weights[name] = d
                shape = (C.c_int64 * max(1, d.dim()))(*d.shape)
                N.check(
                    L.dpt_set_weight(handle, name.encode(), C.c_void_p(d.data_ptr()), shape, d.dim(), code),
                    handle, f"dpt_set_weight({name})",
                )

            for name, (t, kind) in self._packed_cpu.items():
                if kind in ("half", "half_colsum"):
                    d, code = t.to(device=self._device, dtype=self._dtype), _TORCH_TO_DPT[self._dtype]
                    if kind == "half_colsum":  # LayerNorm folded into this GEMM: column sums of the ROUNDED weights
                        assert name.endswith(".w")
                        set_weight(name[:-2] + ".s", d.to(torch.float32).sum(dim=1).contiguous(), N.DPT_F32)
                elif kind == "f32":
                    d, code = t.to(device=self._device, dtype=torch.float32), N.DPT_F32
                else:  # "host": tiny fp32 vectors the library copies into the handle
                    d, code = t.to(dtype=torch.float32).contiguous(), N.DPT_F32
                set_weight(name, d, code)
        self._workspace, self._ws_key, self._io, self._io_extra, self._copy_streams = None, None, {}, None, None

    def _release(self):
        if self._handle is not None:
            N.lib().dpt_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------- plumbing

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _get_workspace(self, B: int, H: int, W: int) -> torch.Tensor:
        key = (B, H, W)
        if self._ws_key != key:
            need = C.c_size_t()
            with torch.cuda.device(self._device):
                N.check(N.lib().dpt_workspace_bytes(self._handle, B, H, W, C.byref(need)), self._handle, "dpt_workspace_bytes")
            self._workspace = None
            self._workspace = torch.empty(int(need.value), dtype=torch.uint8, device=self._device)
            self._ws_key = key
        return self._workspace

    def _check_image(self, x: torch.Tensor):
        device, dtype = self._require_ready()
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an image tensor of shape Bx3xHxW")
        if x.device != device:
            raise RuntimeError(f"Device mismatch! Image: {x.device}, model: {device}")
        if x.dtype != dtype:
            raise RuntimeError(f"Data type mismatch! Image: {x.dtype}, model: {dtype}")
        B, _, H, W = x.shape
        self._check_size(H, W)
        return B, H, W

    def _check_size(self, H: int, W: int):
        mult = (8 if self.model_type == "swinv2" else 2) * self.config["patch_size_px"]
        if H % mult or W % mult:
            raise ValueError(
                f"image size {H}x{W} is not usable: height and width must be multiples of {mult} "
                "(the reference fails inside fusion otherwise)"
            )

    # ------------------------------------------------------------------------------------------------- forward

    def forward(self, image_rgb_normalized_bchw: torch.Tensor) -> torch.Tensor:
        """DPTModel.forward (dpt_model.py:61-83): BxCxHxW -> BxHxW inverse depth, in the model dtype."""
        B, H, W = self._check_image(image_rgb_normalized_bchw)
        img, out = self._io_buffers(B, H, W)
        img.copy_(image_rgb_normalized_bchw)  # fixed buffers keep the recorded launch plan valid across calls
        self.forward_into(img, out)
        return out.clone()

    def forward_into(self, img: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """forward() on caller-owned, contiguous device buffers (no copies; used by bench.py)."""
        B, H, W = self._check_image(img)
        assert img.is_contiguous() and out.is_contiguous() and tuple(out.shape) == (B, H, W) and out.dtype == self._dtype
        ws = self._get_workspace(B, H, W)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_forward(self._handle, C.c_void_p(img.data_ptr()), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), B, H, W, self._stream())
        N.check(rc, self._handle, "dpt_forward")
        return out

    def _check_host_buffers(self, host_img: torch.Tensor, host_out: torch.Tensor):
        device, dtype = self._require_ready()
        for name, t in (("host_img", host_img), ("host_out", host_out)):
            if not isinstance(t, torch.Tensor) or t.device.type != "cpu":
                raise ValueError(f"forward_host: {name} must be a CPU tensor (pinned for full copy speed)")
            if t.dtype != dtype:
                raise RuntimeError(f"Data type mismatch! {name}: {t.dtype}, model: {dtype}")
            if not t.is_contiguous():
                raise ValueError(f"forward_host: {name} must be contiguous")
        if host_img.dim() != 4 or host_img.shape[1] != 3:
            raise ValueError("expected an image tensor of shape Bx3xHxW")
        B, _, H, W = host_img.shape
        self._check_size(H, W)
        if tuple(host_out.shape) != (B, H, W):
            raise ValueError(f"forward_host: host_out must have shape {(B, H, W)}, got {tuple(host_out.shape)}")
        return B, H, W

    def _io_buffers(self, B: int, H: int, W: int, slot: int = 0):
        """device-side (image, depth) pair `slot` for host-buffer forwards of this shape"""
        key = (B, H, W)
        if key not in self._io:
            self._io = {key: (
                torch.empty((B, 3, H, W), dtype=self._dtype, device=self._device),
                torch.empty((B, H, W), dtype=self._dtype, device=self._device),
            )}
            self._io_extra = {}
        if slot == 0:
            return self._io[key]
        extra = getattr(self, "_io_extra", None)
        if extra is None or extra.get("key") != key:
            extra = self._io_extra = {"key": key}
        if slot not in extra:
            extra[slot] = (torch.empty((B, 3, H, W), dtype=self._dtype, device=self._device),
                           torch.empty((B, H, W), dtype=self._dtype, device=self._device))
        return extra[slot]

    def forward_host(self, host_img: torch.Tensor, host_out: torch.Tensor) -> torch.Tensor:
        """Host buffers in/out through dpt_forward_host (H2D + forward + D2H + sync)."""
        B, H, W = self._check_host_buffers(host_img, host_out)
        io = self._io_buffers(B, H, W)
        ws = self._get_workspace(B, H, W)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_forward_host(self._handle, C.c_void_p(host_img.data_ptr()), C.c_void_p(host_out.data_ptr()),
                                          C.c_void_p(io[0].data_ptr()), C.c_void_p(io[1].data_ptr()),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), B, H, W, self._stream())
        N.check(rc, self._handle, "dpt_forward_host")
        return host_out

    def forward_host_async(self, host_img: torch.Tensor, host_out: torch.Tensor, slot: int = 0) -> torch.cuda.Event:
        """Pipelinable host-buffer forward (dpt_forward_host_async): H2D on a copy-in stream, forward on the current
        stream, D2H on a copy-out stream; returns an event to `.synchronize()` on before reading `host_out`. Alternate
        `slot` between 0 and 1 (each slot owns a device image / depth pair) and the copies of one step overlap the
        forward of the other. The caller must not touch `host_img` / `host_out` until the returned event completes."""
        B, H, W = self._check_host_buffers(host_img, host_out)
        io = self._io_buffers(B, H, W, slot)
        ws = self._get_workspace(B, H, W)
        with torch.cuda.device(self._device):
            if getattr(self, "_copy_streams", None) is None:
                self._copy_streams = (torch.cuda.Stream(self._device), torch.cuda.Stream(self._device))
            s_in, s_out = self._copy_streams
            rc = N.lib().dpt_forward_host_async(self._handle, C.c_void_p(host_img.data_ptr()), C.c_void_p(host_out.data_ptr()),
                                                C.c_void_p(io[0].data_ptr()), C.c_void_p(io[1].data_ptr()),
                                                C.c_void_p(ws.data_ptr()), ws.numel(), B, H, W, self._stream(),
                                                C.c_void_p(s_in.cuda_stream), C.c_void_p(s_out.cuda_stream))
            N.check(rc, self._handle, "dpt_forward_host_async")
            done = torch.cuda.Event()
            done.record(s_out)
        return done

    def last_launch_count(self) -> int:
        return int(N.lib().dpt_last_launch_count(self._handle))

    def enable_profiling(self, on: bool = True) -> None:
        """per-launch CUDA-event timing inside the library (dpt_profile_enable)"""
        self._require_ready()
        N.check(N.lib().dpt_profile_enable(self._handle, int(on)), self._handle, "dpt_profile_enable")

    def read_profile(self) -> list[tuple[str, float, float, float]]:
        """[(label, ms, algorithmic flops, algorithmic bytes)] of the last forward; call after a synchronize"""
        L = N.lib()
        out = []
        buf = C.create_string_buffer(128)
        ms, fl, by = C.c_double(), C.c_double(), C.c_double()
        for i in range(L.dpt_profile_count(self._handle)):
            N.check(L.dpt_profile_get(self._handle, i, buf, 128, C.byref(ms), C.byref(fl), C.byref(by)),
                    self._handle, "dpt_profile_get")
            out.append((buf.value.decode(), ms.value, fl.value, by.value))
        return out

    def inference(self, image_bgr, max_side_length=None, use_square_sizing=True) -> torch.Tensor:
        """DPTModel.inference (dpt_model.py:87-109)"""
        with torch.inference_mode():
            x = self.patch_embed.prepare_image(image_bgr, max_side_length, use_square_sizing)
            return self(x)

    def prepare_image_bgr(self, image_bgr, max_side_length=None, use_square_sizing=True, interpolation_mode="bilinear"):
        """DPTModel.prepare_image_bgr (dpt_model.py:113-129)"""
        return self.patch_embed.prepare_image(image_bgr, max_side_length, use_square_sizing, interpolation_mode)

    def verify_input(self, image_rgb_normalized_bchw) -> bool:
        """DPTModel.verify_input (dpt_model.py:133-166): AssertionError on bad input, True otherwise"""
        assert isinstance(image_rgb_normalized_bchw, torch.Tensor), "Image must be provided as a tensor!"
        device, dtype = self._require_ready()
        x = image_rgb_normalized_bchw
        assert x.device == device, f"Device mismatch! Image: {x.device}, model: {device}"
        assert x.dtype == dtype, f"Data type mismatch! Image: {x.dtype}, model: {dtype}"
        shape_str = "x".join(str(s) for s in x.shape)
        assert x.dim() == 4, f"Bad image shape! Image ({shape_str}) should have a shape of BxCXHxW"
        return self.patch_embed.verify_input(x)

    # ------------------------------------------------------------------------------------------------- stages

    def _grid(self, H, W):
        p = self.config["patch_size_px"]
        return H // p, W // p

    def _stage_ws(self, B, gh, gw):
        p = self.config["patch_size_px"]
        return self._get_workspace(B, gh * p, gw * p)

    def _stage_patch_embed(self, image_bchw: torch.Tensor):
        B, H, W = self._check_image(image_bchw)
        gh, gw = self._grid(H, W)
        F = self.config["features_per_token"]
        img = image_bchw.contiguous()
        tokens = torch.empty((B, gh * gw, F), dtype=self._dtype, device=self._device)
        ws = self._get_workspace(B, H, W)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_patch_embed(self._handle, C.c_void_p(img.data_ptr()), C.c_void_p(tokens.data_ptr()),
                                         C.c_void_p(ws.data_ptr()), ws.numel(), B, H, W, self._stream())
        N.check(rc, self._handle, "dpt_patch_embed")
        return tokens, (gh, gw)

    def _stage_encoder(self, patch_tokens: torch.Tensor, patch_grid_hw, capture=None):
        """capture = (probs, block_outs): per-block output tensors or None (module_tree.ImageEncoder, debug only)"""
        self._require_ready()
        gh, gw = int(patch_grid_hw[0]), int(patch_grid_hw[1])
        B, Np, F = patch_tokens.shape
        assert Np == gh * gw and patch_tokens.dtype == self._dtype
        tok = patch_tokens.contiguous()
        if self.model_type == "swinv2":  # hierarchical taps, no cls token (v31_swinv2/image_encoder_model.py:77-98)
            taps = [torch.empty((B, (gh >> k) * (gw >> k), F << k), dtype=self._dtype, device=self._device) for k in range(4)]
        else:
            taps = [torch.empty((B, Np + 1, F), dtype=self._dtype, device=self._device) for _ in range(4)]
        ws = self._stage_ws(B, gh, gw)
        with torch.cuda.device(self._device):
            if capture is None:
                rc = N.lib().dpt_encoder(self._handle, C.c_void_p(tok.data_ptr()), C.byref(N.ptr4(taps)),
                                         C.c_void_p(ws.data_ptr()), ws.numel(), B, gh, gw, self._stream())
            else:
                probs, outs = capture
                n = len(probs)
                p_arr = (C.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in probs])
                o_arr = (C.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in outs])
                rc = N.lib().dpt_encoder_capture(self._handle, C.c_void_p(tok.data_ptr()), C.byref(N.ptr4(taps)), p_arr, o_arr,
                                                 n, C.c_void_p(ws.data_ptr()), ws.numel(), B, gh, gw, self._stream())
        N.check(rc, self._handle, "dpt_encoder" if capture is None else "dpt_encoder_capture")
        return tuple(taps)

    def _nhwc_empty(self, B, Cc, H, W):
        # logical [B,C,H,W] with channels_last strides == the kernels' [B,H,W,C]
        return torch.empty((B, H, W, Cc), dtype=self._dtype, device=self._device).permute(0, 3, 1, 2)

    @staticmethod
    def _as_nhwc(t: torch.Tensor) -> torch.Tensor:
        return t.contiguous(memory_format=torch.channels_last)

    def _stage_reassemble(self, s1, s2, s3, s4, patch_grid_hw):
        gh, gw = int(patch_grid_hw[0]), int(patch_grid_hw[1])
        B = s1.shape[0]
        Cc = self.config["fusion_channels"]
        taps = [t.contiguous() for t in (s1, s2, s3, s4)]
        k0 = 1 if self.model_type == "swinv2" else 4
        sizes = [((gh * k0) >> k, (gw * k0) >> k) for k in range(4)]
        maps = [self._nhwc_empty(B, Cc, h, w) for h, w in sizes]
        ws = self._stage_ws(B, gh, gw)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_reassemble(self._handle, C.byref(N.ptr4(taps)), C.byref(N.ptr4(maps)),
                                        C.c_void_p(ws.data_ptr()), ws.numel(), B, gh, gw, self._stream())
        N.check(rc, self._handle, "dpt_reassemble")
        return tuple(maps)

    def _stage_fusion(self, r1, r2, r3, r4):
        maps = [self._as_nhwc(t) for t in (r1, r2, r3, r4)]
        B, Cc, h0, w0 = maps[0].shape
        k0 = 1 if self.model_type == "swinv2" else 4
        gh, gw = h0 // k0, w0 // k0
        fused = self._nhwc_empty(B, Cc, h0 * 2, w0 * 2)
        ws = self._stage_ws(B, gh, gw)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_fusion(self._handle, C.byref(N.ptr4(maps)), C.c_void_p(fused.data_ptr()),
                                    C.c_void_p(ws.data_ptr()), ws.numel(), B, gh, gw, self._stream())
        N.check(rc, self._handle, "dpt_fusion")
        return fused

    def _stage_head(self, fused: torch.Tensor):
        x = self._as_nhwc(fused)
        B, Cc, h, w = x.shape
        k0 = 2 if self.model_type == "swinv2" else 8
        gh, gw = h // k0, w // k0
        p = self.config["patch_size_px"]
        depth = torch.empty((B, gh * p, gw * p), dtype=self._dtype, device=self._device)
        ws = self._stage_ws(B, gh, gw)
        with torch.cuda.device(self._device):
            rc = N.lib().dpt_head(self._handle, C.c_void_p(x.data_ptr()), C.c_void_p(depth.data_ptr()),
                                  C.c_void_p(ws.data_ptr()), ws.numel(), B, gh, gw, self._stream())
        N.check(rc, self._handle, "dpt_head")
        return depth
