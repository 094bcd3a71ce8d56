/* dpt_b200.h - C ABI of libdpt_b200.so: the B200 (sm_100a) implementation of muggled_dpt's single-image depth
 * inference hot path (reference: /root/reference/muggled_dpt/dpt_model.py:61-83, DPTModel.forward).
 *
 * The reference has no native interface of its own (it is pure PyTorch); these entry points are what a binding for
 * this path binds. Each one names the reference function it replaces. Conventions:
 *   - every function returns 0 on success, a negative dpt_status otherwise; the message is dpt_last_error(handle)
 *     (dpt_last_error(NULL) for failures of dpt_create);
 *   - all pointers are DEVICE pointers unless the name says host; the caller (PyTorch in the Python host) owns every
 *     buffer; the library allocates no device memory;
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it;
 *   - activations are channels-last: token tensors [B, N, F], image-like tensors [B, H, W, C] (== a torch tensor of
 *     logical shape [B, C, H, W] in torch.channels_last memory format);
 *   - 16-bit tensors are bf16 or fp16 according to dpt_config.dtype; "f32" tensors are float.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef DPT_B200_H
#define DPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpt_model_s* dpt_handle;

enum dpt_status {
  DPT_OK = 0,
  DPT_ERR_INVALID = -1,     /* bad argument / unsupported size (e.g. odd patch grid) */
  DPT_ERR_MISSING = -2,     /* a required weight was never set */
  DPT_ERR_WORKSPACE = -3,   /* workspace too small */
  DPT_ERR_CUDA = -4,        /* CUDA runtime / driver error */
  DPT_ERR_UNSUPPORTED = -5  /* no sm_100 device, or feature not built */
};

enum dpt_dtype { DPT_F16 = 0, DPT_BF16 = 1, DPT_F32 = 2 };
enum dpt_variant { DPT_VARIANT_DINOV2 = 0, DPT_VARIANT_BEIT = 1, DPT_VARIANT_SWINV2 = 2 };

/* Mirrors the reference's config dict (v2_depthanything/state_dict_conversion/config_from_original_state_dict.py:29-41). */
typedef struct dpt_config {
  int variant;             /* dpt_variant */
  int dtype;               /* DPT_F16 or DPT_BF16: storage/MMA input type; accumulation is always fp32 */
  int features_per_token;  /* F */
  int num_heads;           /* F / 64 */
  int num_blocks;
  int reassembly_features[4];
  int fusion_channels;     /* C */
  int patch_size_px;       /* 14 (DINOv2) or 16 (BEiT) */
  int base_grid_h, base_grid_w;
  int is_metric;           /* sigmoid instead of the final ReLU (head_model.py:84) */
  float ln_eps;            /* 1e-6 (misc_helpers.py:202); 1e-5 for SwinV2 */
  /* SwinV2 only (make_swinv2_dpt.py:61-72): features_per_token = stage-0 width, stage s has width << s */
  int heads_per_stage[4];
  int layers_per_stage[4];
  int window_h, window_w;
  int pretrained_window[4]; /* 0 = none */
  /* Depth-Anything V1 (v1_depthanything/image_encoder_model.py:92-103): the four taps are the outputs of the LAST four
   * blocks instead of the last block of each quarter of the encoder */
  int taps_last4;
  /* ViT-G (v2_depthanything/components/misc_helpers.py:125-185): the block MLP is the SwiGLU FFN - fc1 holds the doubled
   * inner Linear [2h, F] (gate half first), fc2 the outer Linear [F, h] */
  int mlp_swiglu;
} dpt_config;

/* lifetime ---------------------------------------------------------------------------------------------------- */
int dpt_create(const dpt_config* cfg, dpt_handle* out);
void dpt_destroy(dpt_handle h);
const char* dpt_last_error(dpt_handle h);
const char* dpt_version(void);
int dpt_config_size(void); /* sizeof(dpt_config) the library was compiled with: lets a binding check its struct */

/* weights: replaces nn.Module.load_state_dict of the five sub-models (make_depthanythingv2_dpt.py:55-59).
 * `name` is a packed-weight name (see muggled_dpt_b200/weights.py); the pointer must stay valid for the handle's
 * life. dtype is a dpt_dtype. */
int dpt_set_weight(dpt_handle h, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype);

/* sizing */
int dpt_workspace_bytes(dpt_handle h, int B, int H, int W, size_t* bytes);

/* whole path: DPTModel.forward (dpt_model.py:61-83). img [B,3,H,W] 16-bit NCHW-contiguous, depth [B,H,W] 16-bit. */
int dpt_forward(dpt_handle h, const void* img_bchw, void* depth_bhw, void* workspace, size_t workspace_bytes, int B,
                int H, int W, void* stream);

/* Same, host buffers in / out (pinned or pageable): H2D copy, forward, D2H copy, all on `stream`, then a stream
 * synchronize. dev_img / dev_depth are caller-provided device staging buffers of the same sizes. */
int dpt_forward_host(dpt_handle h, const void* host_img_bchw, void* host_depth_bhw, void* dev_img, void* dev_depth,
                     void* workspace, size_t workspace_bytes, int B, int H, int W, void* stream);

/* Asynchronous, pipelinable form of dpt_forward_host: enqueues the H2D copy on `copy_in_stream`, the forward on `stream`
 * (after that copy) and the D2H copy on `copy_out_stream` (after the forward) and returns at once. Calls that alternate
 * between two (dev_img, dev_depth) pairs overlap the copies of one step with the forward of the other: the library
 * orders every reuse of a device buffer with events it owns (H2D into dev_img waits for the last forward that read it,
 * a forward into dev_depth waits for the last D2H out of it). The caller synchronises `copy_out_stream` before reading
 * host_depth, and does not touch host_img until then. Host buffers should be pinned. */
int dpt_forward_host_async(dpt_handle h, const void* host_img, void* host_depth, void* dev_img, void* dev_depth,
                           void* workspace, size_t workspace_bytes, int B, int H, int W, void* stream,
                           void* copy_in_stream, void* copy_out_stream);

/* pre / post-processing around the path ------------------------------------------ */
/* PatchEmbed.prepare_image (v2_depthanything/patch_embed.py:103-145): uint8 BGR image [IH, IW, 3] (device) -> RGB,
 * antialiased bilinear resize to OH x OW (F.interpolate(mode="bilinear", antialias=True, align_corners=False)),
 * (v / 255 - mean[c]) * inv_std[c], written as 16-bit NCHW [1, 3, OH, OW]. mean / inv_std: 3 host floats (RGB order). */
int dpt_prepare_image(const uint8_t* bgr_hwc, int IH, int IW, void* out_chw, int OH, int OW, const float* mean_rgb,
                      const float* inv_std_rgb, int dtype, void* stream);
/* demo_helpers/postprocess.py:22-102 in one pass pair: scale_prediction (bilinear, align_corners=False) to OH x OW,
 * normalize_01 over the whole scaled tensor, convert_to_uint8 (truncation). depth [B, H, W] 16-bit -> out [B, OH, OW]
 * uint8; minmax: 2 device floats of scratch (receives min, max of the scaled prediction). */
int dpt_postprocess_u8(const void* depth_bhw, int B, int H, int W, uint8_t* out_u8, int OH, int OW, float* minmax,
                       int dtype, void* stream);

/* multi-GPU: the ONE collective of the path - every rank contributes its [B/G, H, W] shard of
 * depth maps, every rank receives [B, H, W]. `nccl_comm` is the caller's ncclComm_t (one process per GPU); the call
 * enqueues ncclAllGather on `stream` and returns. NCCL is resolved at run time from the process (the library does not
 * link it: a PyTorch host already carries one); DPT_ERR_UNSUPPORTED when no NCCL is loaded or loadable. The Python
 * host normally issues the same collective through torch.distributed (muggled_dpt_b200/distributed.py), which owns the
 * communicator there. */
int dpt_allgather_depth(void* nccl_comm, const void* local_depth, void* global_depth, size_t elems_per_rank, int dtype,
                        void* stream);

/* per-stage entry points (reference contract: simple_examples/internal_features.py:38-44) ------------------------ */
/* PatchEmbed.forward (v2_depthanything/patch_embed.py:77-99): img -> tokens [B, gh*gw, F] 16-bit */
int dpt_patch_embed(dpt_handle h, const void* img_bchw, void* tokens, void* workspace, size_t workspace_bytes, int B,
                    int H, int W, void* stream);
/* DinoV2Model4Stages.forward (image_encoder_model.py:80-94): tokens -> four taps [B, 1+gh*gw, F] 16-bit
 * (SwinV2: SwinV2Model4Stages.forward, taps [B, (gh>>s)*(gw>>s), F<<s], no cls token) */
int dpt_encoder(dpt_handle h, const void* tokens, void* const taps[4], void* workspace, size_t workspace_bytes, int B,
                int gh, int gw, void* stream);
/* ReassembleModel.forward (reassembly_model.py:61-94): taps -> maps [B,4g,4g,C] [B,2g,2g,C] [B,g,g,C] [B,g/2,g/2,C] */
int dpt_reassemble(dpt_handle h, const void* const taps[4], void* const maps[4], void* workspace,
                   size_t workspace_bytes, int B, int gh, int gw, void* stream);
/* FusionModel.forward (fusion_model.py:55-80): maps -> fused [B, 8gh, 8gw, C] */
int dpt_fusion(dpt_handle h, const void* const maps[4], void* fused, void* workspace, size_t workspace_bytes, int B,
               int gh, int gw, void* stream);
/* One fusion block on its own (FusionBlock.forward fusion_model.py:148-154; level 3 = TopMostFusionBlock.forward
 * :113-114, prev_fused = NULL): reasm_map, prev_fused [B, map_h, map_w, C] -> out [B, 2*map_h, 2*map_w, C]. This is
 * what experiments/fusion_scaling.py:330-333 calls as dpt_model.fusion.blocks[i](...). Workspace: the one sized by
 * dpt_workspace_bytes for the image this map belongs to is always large enough. */
int dpt_fusion_block(dpt_handle h, int level, const void* reasm_map, const void* prev_fused, void* out,
                     void* workspace, size_t workspace_bytes, int B, int map_h, int map_w, void* stream);
/* dpt_encoder with debug capture (demo_helpers/model_capture.py:15-61 forward hooks): for encoder block i (running
 * index over all blocks / all SwinV2 stages, i < num_blocks) probs[i], if not NULL, receives the attention
 * probabilities softmax(scale q k^T + bias) as [B, heads, N, N] (SwinV2: [B*windows, heads, A, A]) 16-bit - what the
 * reference's nn.Softmax module outputs (transformer_block.py:132) - and block_out[i], if not NULL, the block's output
 * tokens [B, N, F] 16-bit (TransformerBlock.forward :53-65). Either array may be NULL. Never used by dpt_forward. */
int dpt_encoder_capture(dpt_handle h, const void* tokens, void* const taps[4], void* const* probs,
                        void* const* block_out, int num_blocks, void* workspace, size_t workspace_bytes, int B, int gh,
                        int gw, void* stream);
/* MonocularDepthHead.forward (head_model.py:89-106): fused -> depth [B, P*gh, P*gw] */
int dpt_head(dpt_handle h, const void* fused, void* depth, void* workspace, size_t workspace_bytes, int B, int gh,
             int gw, void* stream);

/* single operators, exposed for kernel-level parity tests ------------------------------------------------------- */
/* out[pix, n] = act(sum_tap sum_c A[b, y+dy, x+dx+xoff, c] * Wt[n, tap*kpad + c] + bias[n]) + add1 + add2
 * A: [B,H,W,C] 16-bit NHWC; Wt: [N, taps*kpad] 16-bit (kpad = roundup(C,64)); taps = 1 or 9;
 * out/add1: [B,H,W,N] 16-bit (out_f32 = 0) or f32 (out_f32 = 1); add2, out_relu: 16-bit or NULL. */
int dpt_op_conv_gemm(const void* A, const void* Wt, const float* bias, void* out, const void* add1, const void* add2,
                     void* out_relu, int B, int H, int W, int C, int N, int taps, int xoff, int act, int out_f32,
                     int dtype, void* stream);
/* O = softmax(scale * Q K^T + bias) V, qkv [B,N,3F] 16-bit (F = heads*head_dim, head_dim 64 or 32), out [B,N,F];
 * bias [bias_wmod, heads, N, bias_ld] 16-bit with bias_ld a multiple of 128 (>= N), table (b % bias_wmod) is used
 * for batch entry b; or NULL */
int dpt_op_attention(const void* qkv, const void* bias, int64_t bias_ld, int bias_wmod, void* out, int B, int N,
                     int heads, int head_dim, float scale, int dtype, void* stream);
/* y = LayerNorm(x) * w + b, x [M,F] f32, y [M,F] 16-bit */
int dpt_op_layernorm(const float* x, const float* w, const float* b, void* y, int64_t M, int F, float eps, int dtype,
                     void* stream);
/* bilinear, align_corners=True, NHWC 16-bit */
int dpt_op_resize_bilinear(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int dtype,
                           void* stream);
const char* dpt_op_last_error(void);

/* Per-launch CUDA-event timing of the launches of the most recent dpt_forward / stage call (bench.py's roofline
 * leg). Enable, run, synchronize the stream, then read entry i: label ("gemm256:blk3.fc1", "attn:blk3.", ...),
 * elapsed ms, algorithmic FLOPs (2*MAC, unpadded) and algorithmic HBM bytes of that launch. */
int dpt_profile_enable(dpt_handle h, int on);
int dpt_profile_count(dpt_handle h);
int dpt_profile_get(dpt_handle h, int i, char* label, int label_len, double* ms, double* flops, double* bytes);

/* number of kernels the most recent dpt_forward / stage call launched (bench.py's gpu_launches) */
int dpt_last_launch_count(dpt_handle h);

#ifdef __cplusplus
}
#endif
#endif /* DPT_B200_H */
