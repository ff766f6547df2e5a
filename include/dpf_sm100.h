/* dpf_sm100.h -- C ABI of libdpf_sm100.so: the B200 (sm_100a) stereo hot path of DualPixelFace.
 *
 * The reference has exactly one native boundary on this path, the pybind11 module `DCN`
 * (src/module/dcn3d/src/vision.cpp:4-7; deform_conv_forward / deform_conv_backward,
 * src/module/dcn3d/src/deform_conv.h:10-29,49-69).  Everything else on the path is stock PyTorch called from
 * src/model/<name>/modules.py.  This header is the C-ABI a maintainer binds instead (ctypes stub in
 * INTEGRATION.md); each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all buffers,
 *     nothing is allocated behind the ABI; 16-byte alignment is required for tensors.
 *   - activations are bf16, channels-last: features [B,H4,W4,C], volumes / 3-D activations [B,D,H,W,C];
 *     head costs, disparities, statistics are fp32.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, no global mutable
 *     state except the per-thread last-error string.
 *   - return 0 on success, non-zero on error (message via dpf_last_error()); no exceptions cross the ABI.
 *   - sm_100a only: dpf_device_check() fails on any other device.  There is no CPU path.
 */
#ifndef DPF_SM100_H
#define DPF_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPF_ABI_VERSION 1

int dpf_abi_version(void);
const char* dpf_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x with >= 200 KB opt-in shared memory. */
int dpf_device_check(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches evidence). */
long long dpf_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * (1) Integer-shift cost volume.  Replaces CostVolume.build_concat_volume / build_gwc_volume
 *     (src/model/psmnet/modules.py:223-262) and the difference volume of src/model/stereonet/mainmodel.py:100-114.
 *     vol[b,i,h,w,:] = f(ref[b,h,w,:], tgt[b,h+s_i,w,:]) if 0 <= h+s_i < H4 else 0, s_i = shifts_host[i].
 *     mode 0 concat (Cv = 2C: ref | tgt), 1 difference (Cv = C: ref - tgt), 2 group-wise correlation
 *     (Cv = G: -mean over C/G channels of ref*tgt).  C % 8 == 0, G | C, D <= 16.
 * ------------------------------------------------------------------------------------------------- */
int dpf_costvol_fwd(int mode, const void* ref, const void* tgt, void* vol, int B, int H4, int W4, int C, int D, int G,
                    const int* shifts_host, void* stream);
/* gradient of the above wrt ref / tgt (gather form, no atomics).  dvol has the layout of vol. */
int dpf_costvol_bwd(int mode, const void* ref, const void* tgt, const void* dvol, void* dref, void* dtgt, int B, int H4,
                    int W4, int C, int D, int G, const int* shifts_host, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (2) ASM sampling + blend (StereoDPNet volume).  Replaces subpixel_shift.forward (src/module/asm/asm.py:87-127)
 *     and the softmax-blend tail of MaskingAttention.forward (asm.py:160-171) + the volume writes of
 *     CostVolume.build_concat_volume (src/model/stereodpnet/modules.py:181-197).
 *
 *     dpf_asm_sample_fwd: out[s,b,h,w,:] = sum_{i,j in {0,1}} rw[s,h,i]*cw[s,w,j] * x[b, ri[s,h,i], ci[s,w,j], :]
 *     (index < 0 contributes 0).  Tables are built on the host with the reference's own op sequence so that
 *     sampling coordinates are bit-exact; layout ri/rw [S,H4,2], ci/cw [S,W4,2] (int32 / fp32, device).
 *
 *     dpf_asm_blend_fwd: y[b,h,w,c] = mean_s( x_s * softmax_s( sigmoid( logit_s * a[b,c] + d[b,c] ) ) ) written to
 *     vol[b, d0..d0+D_rep-1, h, w, ch_off + c] of a [B,D_vol,H4,W4,Cvol] volume (a,d = folded InstanceNorm3d
 *     affine, fp32 [B,C]).  D_rep = D_vol reproduces the reference's cached-first-level behaviour in one pass.
 * ------------------------------------------------------------------------------------------------- */
int dpf_asm_sample_fwd(const void* x, void* out, int B, int H4, int W4, int C, int S, const int* ri, const float* rw,
                       const int* ci, const float* cw, void* stream);
int dpf_asm_blend_fwd(const void* samples, const void* logits, const float* in_a, const float* in_d, void* vol, int B,
                      int H4, int W4, int C, int S, int D_vol, int d0, int D_rep, int ch_off, int Cvol, void* stream);
/* per-(b,c) sum and sum of squares over all positions of x [B,P,C] bf16 -> stats [B,C,2] fp32 (fully written).  Deterministic:
 * two passes over a fixed summation tree, no floating-point atomics; ws = caller-owned scratch of
 * dpf_channel_stats_ws_floats(B, P, C) floats (no hidden allocation). */
long long dpf_channel_stats_ws_floats(int B, long long P, int C);
int dpf_channel_stats(const void* x, float* stats, float* ws, int B, long long P, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (3) 3-D convolution, implicit GEMM on tcgen05 tensor cores (fp32 accumulation in TMEM), with the
 *     per-channel affine (folded BatchNorm / bias), residual add and ReLU fused in the epilogue.
 *     Replaces every nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d + ReLU (+ add) of PSMNetHGAggregation
 *     (src/model/stereodpnet/modules.py:204-337, convbn_3d src/module/asm/basics.py:32-36) and the mask
 *     convolutions of MaskingAttention (src/module/asm/asm.py:141-146).
 *       y = relu?( conv(x, w) * scale[c] + shift[c] + residual )
 *     kind: 0 = 3x3x3 stride 1 pad 1; 1 = 3x3x3 stride 2 pad 1; 2 = transposed 3x3x3 stride 2 pad 1 out_pad 1;
 *           3 = 1x3x3 (per-plane) pad (0,1,1); 4 = 1x1x1.
 *     x [B,D,H,W,Cin] bf16; w packed by dpf_conv3d_weight_elems / the host packer (layout [tap][Cin/8][Npad][8]; kind 2: the
 *     same 27 tap matrices regrouped into 15 operands of 1, 2 or 4 taps [Cin/8][ncls*Npad][8] -- the taps of the output-parity
 *     classes that read the same input shift are one MMA -- in the order of ops.fuse_t2_weight / the table in dpf_conv3d_fwd);
 *     y [B,Do,Ho,Wo,y_cstride] bf16 (or fp32 when y_f32), written at channel offset y_coff;
 *     residual has y's layout (bf16, or fp32 when y_f32).  scale/shift may be NULL (1 / 0).
 *     stats (optional, fp32 [2*Cout], NOT zeroed by the call) accumulates sum / sum-of-squares of the raw
 *     convolution output per channel (training-mode BatchNorm).
 *     Per-launch limits (wider layers are split on the host with y_coff / x_coff): stride-1 kinds Cin=32 -> Cout<=64,
 *     Cin=64 -> Cout<=32; stride-2 Cin=32, Cout<=32; transposed Cin in {32,64}, Cout<=32.
 * ------------------------------------------------------------------------------------------------- */
typedef struct dpf_conv3d_args {
  int kind;
  int B, D, H, W;          /* input grid */
  int Cin, Cout;           /* Cin in {32,64}; Cout in {1..64} */
  const void* x;
  const void* w;
  void* y;
  int y_f32, y_cstride, y_coff;
  const float* scale;
  const float* shift;
  const void* residual;
  int relu;
  float* stats;
  int x_cstride, x_coff;   /* channel stride / offset of x (0,0 = dense Cin); lets a launch read a channel window */
  int res_pre;             /* 1: y = relu?((conv + residual) * scale + shift)  (K-split partial sums) */
  float slope;             /* relu != 0: v > 0 ? v : slope * v (0 = ReLU; LeakyReLU(0.2) of StereoNet's filter, src/model/stereonet/
                              mainmodel.py:43-49); non-zero only for kind 0 with Cout <= 32 */
} dpf_conv3d_args;
int dpf_conv3d_fwd(const dpf_conv3d_args* args, void* stream);
/* number of bf16 elements of the packed weight buffer for (kind, Cin, Cout) */
long long dpf_conv3d_weight_elems(int kind, int Cin, int Cout);

/* ---------------------------------------------------------------------------------------------------
 * (4) Fused trilinear x4 upsample (align_corners) + softmax over 4*D bins + soft-argmin.
 *     Replaces F.interpolate(..., 'trilinear') (src/model/stereodpnet/modules.py:327-334) + disp_regression.forward
 *     (modules.py:352-362) without materialising the [B,4D,H,W] tensor.
 *     cost [B,D,H4,W4] fp32 -> disp [B,4*H4,4*W4] fp32; prob (optional, may be NULL) [B,4D,4*H4,4*W4] fp32.
 *     bin k has value mindisp + k*step.
 * ------------------------------------------------------------------------------------------------- */
int dpf_regress_fwd(const float* cost, float* disp, float* prob, int B, int D, int H4, int W4, float mindisp, float step,
                    void* stream);
/* dpf_regress_fwd with HALF-PIXEL (align_corners=False) coordinates on all three axes: NNet's F.interpolate(scale_factor=4,
 * 'trilinear', align_corners=False) + disp_regression (src/model/nnet/mainmodel.py:149-152, src/model/nnet/modules.py:191-217).
 * D = 8 only. */
int dpf_regress_fwd_halfpixel(const float* cost, float* disp, float* prob, int B, int D, int H4, int W4, float mindisp, float step,
                              void* stream);
/* Soft-argmin WITHOUT up-sampling (StereoNet's disp_regression, src/model/stereonet/modules.py:99-120): cost [B,D,P] fp32 ->
 * disp [B,P] = sum_d softmax_d(cost) * (mindisp + d*step); prob (optional, may be NULL) [B,D,P] = the softmax.  D <= 64. */
int dpf_softargmin_fwd(const float* cost, float* disp, float* prob, int B, int D, long long P, float mindisp, float step, void* stream);
/* dcost [B,D,H4,W4] (zeroed by the call) from ddisp [B,H,W]; recomputes the softmax. */
int dpf_regress_bwd(const float* cost, const float* ddisp, float* dcost, int B, int D, int H4, int W4, float mindisp,
                    float step, void* stream);
/* dpf_regress_fwd on a ROW TILE (BASELINE config 5): cost [B,D,H4loc,W4] holds the quarter-resolution rows q_row0 ..
 * q_row0+H4loc-1 of an image H4glob rows tall (q_row0 may be negative / the tile may extend past the image: halo rows), disp
 * [B,Hout,4*W4] (and prob [B,4D,Hout,4*W4]) the full-resolution rows y_row0 .. y_row0+Hout-1.  Source coordinates use the GLOBAL
 * align_corners scale, so tiles reproduce the untiled result exactly; the call fails if the tile does not cover the rows it needs. */
int dpf_regress_fwd_tile(const float* cost, float* disp, float* prob, int B, int D, int H4loc, int W4, int H4glob, int q_row0,
                         int Hout, int y_row0, float mindisp, float step, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (5) ANM front end.  Replaces the nearest x0.25 down-sample, sample_with_sort and grid_maker_3d of
 *     src/model/stereodpnet/normal_module.py:120-138,80-118,156-167.
 *     dpf_anm_select: per quarter-res pixel take d = disp[b,4h,4w]*0.25, pick the K levels nearest to d
 *       (ties -> lower level index; documented rule), sorted ascending -> idx [B,K,H4,W4] int32,
 *       ray-scaled depth coordinates coord [B,K,H4,W4,3] fp32 (un-normalised) and per-sample min/max [B,2] fp32
 *       (minmax must be initialised to +inf/-inf by the caller).
 *     dpf_anm_gather: fv[b,k,h,w,0:C] = out3[b,idx,h,w,:], fv[..., C:C+3] = (coord-min)/(max-min+1e-6),
 *       fv[..., C+3:Cpad] = 0.   out3 [B,D,H4,W4,C] bf16; fv [B,K,H4,W4,Cpad] bf16.
 * ------------------------------------------------------------------------------------------------- */
int dpf_anm_select(const float* disp, const float* kinv, const float* abvalue, const float* levels_host, int* idx,
                   float* coord, float* minmax, int B, int D, int K, int H4, int W4, void* stream);
int dpf_anm_gather(const void* out3, const int* idx, const float* coord, const float* minmax, void* fv, int B, int D,
                   int K, int H4, int W4, int C, int Cpad, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (6) 3-D deformable convolution (D3D), 3x3x3 stride 1 pad 1, groups 1.  Replaces DCN.deform_conv_forward
 *     (src/module/dcn3d/src/deform_conv.h:10-29 -> src/cuda/deform_conv_cuda.cu:18-126 and the im2col kernel
 *     src/cuda/deform_im2col_cuda.cuh:192-265) without the [27*Cin, B*D*H*W] column buffer.
 *     x [B,D,H,W,x_cstride] bf16 (first Cin_pad in {48,64} channels used), offset [B,D,H,W,off_cstride>=81] fp32
 *     ((d,h,w) per tap; a stride of 96 lets the offset conv use 128-bit fp32 stores), w packed like kind 0 with Cin_pad,
 *     y = relu?( dconv * scale + shift ) -> [B,D,H,W,Cout] bf16.
 * ------------------------------------------------------------------------------------------------- */
int dpf_dcn3d_fwd(const void* x, const float* offset, const void* w, const float* scale, const float* shift, void* y,
                  int B, int D, int H, int W, int Cin_pad, int x_cstride, int off_cstride, int Cout, int relu, void* stream);

/* (6b) D3D backward.  Replaces DCN.deform_conv_backward (src/module/dcn3d/src/cuda/deform_conv_cuda.cu:128-285; col2im and
 *      col2im_coord kernels src/cuda/deform_im2col_cuda.cuh:267-405; called from functions/deform_conv_func.py:42-60).
 *      x [B,D,H,W,x_cstride] bf16 (first 64 channels used, zero padded), offset and doffset [.,off_cstride] fp32 (81 used),
 *      dy [B,D,H,W,64] bf16.
 *      bwd_data:   w_t = W^T packed like kind 0 with the roles of the channels swapped ([27][64/8 o][64 c][8]);
 *                  dx [B,D,H,W,x_cstride] fp32 is ACCUMULATED into (zero it first; dx_channels = 32 restricts the scatter to
 *                  channels [0,32) when the caller needs no gradient for the rest), doffset [.,81] fp32 is written.
 *      bwd_weight: dw [27][64 c][64 o] fp32 is ACCUMULATED into (zero it first). */
int dpf_dcn3d_bwd_data(const void* x, const float* offset, const void* dy, const void* w_t, float* dx, float* doffset,
                       int B, int D, int H, int W, int x_cstride, int off_cstride, int dx_channels, void* stream);
int dpf_dcn3d_bwd_weight(const void* x, const float* offset, const void* dy, float* dw, int B, int D, int H, int W,
                         int x_cstride, int off_cstride, void* stream);

/* (6c) Backward of the memory-bound ANM ops (normal_module.py:140-194).
 *      tail_bwd:   x [B*K,H4,W4,3] bf16 (the forward input), dout [B,3,4*H4,4*W4] fp32 -> dx [B*K,H4,W4,3] fp32, ACCUMULATED
 *                  (zero it first): adjoint of mean_k(sigmoid(bilinear x4, align_corners=True)) * 2 - 1 (normal_module.py:186-192).
 *      gather_bwd: dfv_f32 and/or dfv_bf16 [B,K,H4,W4,Cpad] (summed; either may be NULL), idx [B,K,H4,W4] ->
 *                  dout3 [B,D,H4,W4,C] bf16, fully written (levels not selected get zero); adjoint of dpf_anm_gather. */
int dpf_anm_tail_bwd(const void* x, const float* dout, float* dx, int B, int K, int H4, int W4, void* stream);
int dpf_anm_gather_bwd(const float* dfv_f32, const void* dfv_bf16, const int* idx, void* dout3, int B, int D, int K,
                       int H4, int W4, int C, int Cpad, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (7) Fused per-channel bias + residual + activation on channels-last bf16 (one memory pass).  Replaces the
 *     aten::add_ (conv bias) / aten::add (skip) / aten::prelu|relu|leaky_relu chains around the cuDNN 2-D convolutions of
 *     the encoders (src/model/stereodpnet/modules.py:21-134) and of ANM's n_convs (normal_module.py:59-66); writing into
 *     a channel window of a wider tensor also replaces torch.cat of the DPBlock dilated branches (modules.py:42-44).
 *       y[p, y_coff + c] = act(x[p,c] + bias[c] + res[p,c]),  act(v) = v > 0 ? v : slope*v  (0 = ReLU, 1 = identity)
 *     x, res [npix, C] bf16; y [npix, y_cstride] bf16; bias fp32 [C]; bias / res may be NULL.
 * ------------------------------------------------------------------------------------------------- */
int dpf_bias_act(const void* x, const float* bias, const void* res, void* y, long long npix, int C, int y_cstride,
                 int y_coff, float slope, void* stream);

/* Feature-pyramid tail of the StereoDPNet encoder (src/model/stereodpnet/modules.py:128-133): one pass writes
 * out[N,h,w,3C] = cat(f1[N,h,w,C], bilinear(f2[N,h2,w2,C]), bilinear(f3[N,h3,w3,C])), align_corners=True, bf16 channels-last. */
int dpf_pyramid_cat(const void* f1, const void* f2, const void* f3, void* out, int N, int h, int w, int h2, int w2, int h3,
                    int w3, int C, void* stream);
/* dpf_pyramid_cat on a ROW CROP: the three maps hold the rows row0.. / row0/2.. / row0/4.. of levels that are hglob, hglob/2,
 * hglob/4 rows tall; source rows follow the GLOBAL align_corners scale (BASELINE config 5: the encoder runs on this rank's rows
 * plus its receptive-field margin). */
int dpf_pyramid_cat_tile(const void* f1, const void* f2, const void* f3, void* out, int N, int h, int w, int h2, int w2, int h3,
                         int w3, int C, int hglob, int row0, void* stream);
/* FPN top-down merge (torchvision FeaturePyramidNetwork.forward, used at src/model/stereodpnet/modules.py:83,124):
 * y[N,h,w,C] = x + bias + nearest_upsample(top[N,ht,wt,C]); bf16 channels-last, one pass. */
int dpf_fpn_merge(const void* x, const float* bias, const void* top, void* y, int N, int h, int w, int ht, int wt, int C,
                  void* stream);
/* ref_feature = ref_fea.max(1)[0] (src/model/stereodpnet/mainmodel.py:104): per-pixel maximum over the C channels of a
 * channels-last bf16 map x[npix][C] -> y[npix] fp32. */
int dpf_channel_max(const void* x, float* y, long long npix, int C, void* stream);

/* 2-D 3x3 convolution (stride 1, dilation dil in {1,3,5}, pad = dil) + per-channel affine + residual + ReLU / LeakyReLU on channels-last bf16
 * images, on the kd-fused tcgen05 kernel (the image's kh taps are the fused taps; see conv3d_tc.cu).  Replaces the
 * nn.Conv2d + BatchNorm2d (folded) + PReLU / ReLU triples of the StereoDPNet encoder blocks
 * (src/model/stereodpnet/modules.py:21-54, convbn of src/module/asm/basics.py:17-22) for Cin = 32:
 *   y[n,h,w, y_coff+co] = act( conv(x[..., x_coff:x_coff+32]; w)[co] * scale[co] + shift[co] + residual[n,h,w, y_coff+co] )
 * w = dpf_conv3d weight packing of the [Cout,32,3,3,3] tensor whose (kd, 1, kw) taps hold the 2-D kernel (kh -> kd).
 * Cout <= 32 per launch; act(v) = relu ? (v > 0 ? v : slope*v) : v. */
int dpf_conv2d_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, const void* residual,
                   int N, int H, int W, int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff,
                   int dil, int relu, float slope, void* stream);

/* Stride-2 3x3x3 convolution (pad 1) + per-channel affine + ReLU, plane-streamed on tcgen05 (conv3d_s2.cu): every input plane is
 * staged once per tile, N = all output channels (<= 64 for Cin = 32, <= 32 for Cin = 64), input channels in 32-channel windows that
 * accumulate in TMEM.  Replaces dresK.conv1 / conv3 (nn.Conv3d k3 s2 p1 + BatchNorm3d + ReLU, src/model/stereodpnet/modules.py:
 * 208,213) in ONE launch where dpf_conv3d_fwd kind 1 needed 2 / 4.  x [B,D,H,W,x_cstride] -> y [B,ceil(D/2),ceil(H/2),ceil(W/2),
 * y_cstride] bf16; w = dpf_conv3d_fwd's packing of the [Cout,Cin,3,3,3] tensor ([27][Cin/8][Npad][8]).  Deterministic. */
int dpf_conv3d_s2_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, int B, int D, int H, int W,
                      int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff, int relu, void* stream);

/* 32 -> 1 channel 3x3x3 convolution (stride 1, pad 1), fp32 output: the `nn.Conv3d(32, 1, 3, 1, 1)` closing each classifK of the
 * hourglass (src/model/stereodpnet/modules.py:288-296, cumulative adds :323-325) and StereoNet's conv3d_alone
 * (src/model/stereonet/mainmodel.py:50-51).  Bandwidth-bound formulation (conv3d_head.cu): the 27 taps are the N dimension of ONE
 * MMA pair per 128 positions, the partial planes are combined with their shifts on the CUDA cores.
 *   y[b,d,h,w] = sum_{c,kd,kh,kw} w[c,kd,kh,kw] * x[b,d+kd-1,h+kh-1,w+kw-1,c] + shift + residual[b,d,h,w]
 * x [B,D,H,W,x_cstride] bf16 (channels 0..31 are used); w bf16 [4 (c/8)][32 (tap, 27 real + 5 zero rows)][8 (c%8)];
 * y / residual [B,D,H,W] fp32 (residual may be NULL, may alias y).  Deterministic. */
int dpf_conv3d_head_fwd(const void* x, const void* w, float* y, const float* residual, float shift, int B, int D, int H, int W,
                        int x_cstride, void* stream);

/* Encoder stem: 3x3 / stride 2 / pad 1 convolution from an 8-channel (3 real + zero padding) channels-last bf16 image to 32
 * channels + bias + ReLU: `convbn(input_channel, 32, 3, 2, 1, 1)` + ReLU with the BatchNorm folded, the first layer of both
 * encoders (src/model/stereodpnet/modules.py:66-68, src/model/psmnet/modules.py:72-74).  Bandwidth-bound (K = 27): warp-level
 * mma.sync on a shared-memory window, coalesced 64-byte pixel rows out (stem_conv.cu).
 * x [N,H,W,8] bf16; w [80][32] bf16 row-major with k = (kh*3 + kw)*8 + ci (rows 72..79 and ci >= Cin are zero);
 * bias fp32 [32] or NULL; y [N,ceil(H/2),ceil(W/2),32] bf16. */
int dpf_stem_conv_fwd(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int relu, void* stream);

/* 2-D 3x3 convolution, stride 1, ANY dilation (pad = dil), Cin in {32, 64, 96}, Cout <= 96 in ONE launch, on a dedicated tcgen05
 * implicit-GEMM kernel (conv2d_tc.cu: dilation by residue-class sub-images, input channels consumed in 32 / 48-channel windows
 * that accumulate in TMEM, N = Cout).  Replaces the six bias-free Conv2d + LeakyReLU(0.1) `convtext` layers of ANM
 * (src/model/stereodpnet/normal_module.py:14-19,59-66: 64->96->96->64->64->32->3, dilation 1,2,4,8,1,1):
 *   y[n,h,w, y_coff+co] = act( conv(x[..., x_coff:x_coff+Cin]; w)[co] * scale[co] + shift[co] + residual[n,h,w, y_coff+co] )
 * x [N,H,W,x_cstride], y / residual [N,H,W,y_cstride] bf16 channels-last; w bf16 [9 taps (kh,kw)][Cin/8][Npad][8] with
 * Npad = dpf_conv2d_tc_npad(Cout) (output channels zero-padded); ceil8(Cout) channels are written (the extra ones are 0).
 * scale / shift fp32 [Cout] or NULL; act(v) = (relu & 1) ? (v > 0 ? v : slope*v) : v.  relu & 2: the residual is added AFTER the
 * activation, y = act(conv * scale + shift) + residual (BasicBlock of StereoNet, src/model/stereonet/modules.py:19-29).
 * Deterministic (single MMA issuer). */
int dpf_conv2d_tc_npad(int Cout);
long long dpf_conv2d_tc_weight_elems(int Cin, int Cout);
int dpf_conv2d_tc_fwd(const void* x, const void* w, void* y, const float* scale, const float* shift, const void* residual,
                      int N, int H, int W, int Cin, int Cout, int x_cstride, int x_coff, int y_cstride, int y_coff,
                      int dil, int relu, float slope, void* stream);

/* ANM tail: bilinear x4 upsample (align_corners) -> sigmoid -> mean over K -> *2-1 in one pass.  Replaces final_layer and
 * the mean / rescale of ANM.forward (src/model/stereodpnet/normal_module.py:69-72,185-190).
 * x [B*K,H4,W4,3] bf16 (channels-last) -> out [B,3,4*H4,4*W4] fp32. */
int dpf_anm_tail(const void* x, float* out, int B, int K, int H4, int W4, void* stream);
/* Same op on a ROW TILE of the image and with a channel pitch: x [B*K, H4loc, W4, x_cstride] (3 real channels) holds the
 * quarter-resolution rows q_row0 .. q_row0+H4loc-1 of an image that is H4glob rows tall; out [B,3,Hout,4*W4] holds the
 * full-resolution rows y_row0 .. y_row0+Hout-1.  Source coordinates use the GLOBAL align_corners scale (H4glob-1)/(4*H4glob-1),
 * so the tiles of a row-split image (BASELINE config 5) reproduce the untiled result exactly.  dpf_anm_tail == the call with
 * x_cstride 3, H4glob = H4loc, q_row0 = y_row0 = 0, Hout = 4*H4loc. */
int dpf_anm_tail_tile(const void* x, float* out, int B, int K, int H4loc, int W4, int x_cstride, int H4glob, int q_row0,
                      int Hout, int y_row0, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (8) Training-mode BatchNorm3d around the convolution (batch statistics; convbn_3d of src/module/asm/basics.py:32-36).
 *     With z the raw conv output [npix, C] bf16 and statistics from dpf_channel_stats:
 *       dpf_affine_act   : y = act(z*scale[c] + bias[c] + res)                       (forward normalise + residual + ReLU)
 *       dpf_bn_bwd_reduce: sums[0:C] = sum g, sums[C:2C] = sum g*z, g = dy * (y > 0 ? 1 : slope) (relu; slope 0 = ReLU,
 *                          0.2 = the LeakyReLU of StereoNet's 3-D filter) or dy                  (zeroes sums itself)
 *       dpf_bn_bwd_apply : dz = A*(g - K1 - (z - MU)*K2) with coef = [A | K1 | K2 | MU] (fp32 [4C]); dres = g if non-NULL
 * ------------------------------------------------------------------------------------------------- */
int dpf_affine_act(const void* x, const float* scale, const float* bias, const void* res, void* y, long long npix, int C,
                   float slope, void* stream);
/* Per-channel coefficient arithmetic of the train-mode BatchNorm in one launch each (replaces ~25 tiny element-wise launches per
 * layer and step): dpf_bn_fwd_coefs: stats [C][2] (sum z, sum z^2 over n elements) -> out [4][C] = a | b | mean | inv_std and the
 * running statistics (NULL: not tracked) updated in place with `momentum` (unbiased variance), as nn.BatchNorm does;
 * dpf_bn_bwd_coefs: sums [2][C] (dpf_bn_bwd_reduce) + the forward's out -> dgamma, dbeta [C] and coef [4][C] for dpf_bn_bwd_apply. */
int dpf_bn_fwd_coefs(const float* stats, const float* gamma, const float* beta, float n, float eps, float momentum,
                     float* running_mean, float* running_var, float* out, int C, void* stream);
int dpf_bn_bwd_coefs(const float* sums, const float* fwd, float n, float* dgamma, float* dbeta, float* coef, int C, void* stream);
int dpf_bn_bwd_reduce(const void* dy, const void* y, const void* z, float* sums, long long npix, int C, int relu, float slope,
                      void* stream);
int dpf_bn_bwd_apply(const void* dy, const void* y, const void* z, const float* coef, void* dz, void* dres, long long npix,
                     int C, int relu, float slope, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (9) Backward of the ASM volume pieces (autograd of src/module/asm/asm.py:87-127,160-171 in the reference).
 *     dpf_asm_blend_bwd : dy = sum over the D_rep slices of dvol[b,d0..,h,w,ch_off+c]; outputs the direct-path gradient
 *                         wrt the samples (dsamples) and the gradient wrt the normalised logits (dlhat), both [B,S,H4,W4,C] bf16.
 *     dpf_asm_sample_bwd: dfeat [B,H4,W4,C] fp32 (zeroed by the call) += transpose of the table gather of dpf_asm_sample_fwd.
 * ------------------------------------------------------------------------------------------------- */
int dpf_asm_blend_bwd(const void* samples, const void* logits, const float* in_a, const float* in_d, const void* dvol,
                      void* dsamples, void* dlhat, int B, int H4, int W4, int C, int S, int D_vol, int d0, int D_rep, int ch_off,
                      int Cvol, void* stream);
int dpf_asm_sample_bwd(const void* dsamples, float* dfeat, int B, int H4, int W4, int C, int S, const int* ri, const float* rw,
                       const int* ci, const float* cw, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (10) Weight gradient of the stride-1 convolution kinds (0, 3, 4) on tcgen05 (voxel positions are the GEMM K dimension,
 *      both operands MN-major straight from the forward kernel's shared-memory staging).  Replaces the dW half of autograd
 *      through nn.Conv3d (src/model/stereodpnet/modules.py:204-337).
 *        dw[tap][ci][co] += sum_{b,d,h,w} x[b, (d,h,w)+tap-pad, x_coff+ci] * dz[b,d,h,w, z_coff+co]
 *      x [B,D,H,W,x_cstride], dz [B,D,H,W,z_cstride] bf16 (dz channels padded to a multiple of 8); dw fp32
 *      [ntaps][Cin][Cout], accumulated with atomics (zero it first).  Cin in {32,64}, Cout <= 32 per launch.
 * ------------------------------------------------------------------------------------------------- */
int dpf_conv3d_wgrad(int kind, const void* x, const void* dz, float* dw, int B, int D, int H, int W, int Cin, int x_cstride,
                     int x_coff, int Cout, int z_cstride, int z_coff, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * (11) Fused training losses: masked smooth-L1 over the n disparity heads + the element-wise cosine normal
 *      loss, value and gradient in one pass.  Replaces SMOOTHL1Loss.forward ('given' conversion, disparity target;
 *      src/loss/depth/smoothL1.py:15-49) and COSINELoss.forward (masked branch; src/loss/normal/cosine.py:35-55).
 *        pred_depth [B,n,H,W], gt_disp / mask [B,H,W] (mask > 0 selects; NULL = all), pred_normal / gt_normal [B,3,H,W]
 *        (pred_normal NULL = depth only), all fp32.
 *        sums [n+2] = { sum_mask sl1(pred_i - gt) (i < n), sum mask, sum_mask sum_c (1 - sim_c) }  (written, deterministic)
 *        g_depth [B,n,H,W] = mask * d sl1 / d pred_i;   g_normal [B,3,H,W] = mask * d sum_c(1 - sim_c) / d pred_normal
 *      The caller applies the scalars: smoothL1 = sum_i w_i sums[i] / sums[n]; cosine = sums[n+1] / (3 sums[n]).
 *      ws: caller-owned scratch of dpf_fused_losses_ws_floats(B*H*W) floats.
 * ------------------------------------------------------------------------------------------------- */
long long dpf_fused_losses_ws_floats(long long npix);
int dpf_fused_losses(const float* pred_depth, int n_heads, const float* gt_disp, const float* mask, const float* pred_normal,
                     const float* gt_normal, float* g_depth, float* g_normal, float* ws, float* sums, int B, int H, int W,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPF_SM100_H */
