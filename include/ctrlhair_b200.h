/*
 * ctrlhair_b200 — C ABI of the B200 (sm_100a) SEAN/SPADE generator hot path.
 *
 * The reference (XuyangGuo/CtrlHair) has no FFI layer: its boundary is the Python call surface
 *   sean_codes/models/pix2pix_model.py:39-74   Pix2PixModel.forward(data, mode)
 *   sean_codes/models/networks/generator.py:72-109  SPADEGenerator.forward(input, rgb_img, obj_dic)
 *   hair_editor.py:159-179                      HairEditor.gen_img(code, parsing)
 * The entry points below are what a ctypes stub behind those methods binds.
 * Plain pointers and sizes only; all device pointers are CUDA device memory on the current device,
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success and a negative
 * code on failure; chb_last_error() returns a human readable message for the calling thread.
 * No function allocates device memory: workspaces are sized by the library and owned by the caller.
 */
#ifndef CTRLHAIR_B200_H
#define CTRLHAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_OK 0
#define CHB_ERR_ARG (-1)
#define CHB_ERR_CUDA (-2)
#define CHB_ERR_ARCH (-3)

int chb_version(void);
const char* chb_last_error(void);
/* 0 when the current device is compute capability 10.x, CHB_ERR_ARCH otherwise. */
int chb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * Operator 1: implicit-GEMM convolution (the tcgen05 kernel every conv of the path runs on).
 * Replaces nn.Conv2d / F.conv2d call sites of normalization.py:172-173,241-256,
 * architecture.py:35-45,75,79,90 and generator.py:33,51.
 *
 * out[b,y,x,n] = epilogue( sum_seg sum_tap sum_c  A_seg[b, y+dy, x+dx, ch_off+c] * W_seg[(b,) n, tap*C+c] )
 * A operands are fp16 NHWC, weights fp16 [rows][K] (K-major), accumulation fp32.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a;      /* fp16 activations, logical [B,H,W,Ca] with element strides below        */
  int64_t a_sb, a_sy, a_sx; /* element strides of image / row / pixel (channel stride is 1)      */
  int Ca;             /* channels present in the tensor                                          */
  int ch_off;         /* first channel this segment reads                                        */
  int C;              /* channels read: 32, or a multiple of 64                                  */
  int taps;           /* 9 = 3x3 zero-pad 1, 1 = 1x1                                             */
  const void* w;      /* fp16 weights [(B,) Nrows, taps*C]; k = tap*C + c, tap = ky*3+kx         */
  int per_image;      /* 1: weights carry a leading image dimension                              */
  int64_t w_sb;       /* per_image: element stride between images (0 = dense Nrows*taps*C)       */
  int a_pad;          /* the A tensor carries an explicit border of a_pad pixels per side (e.g. a reflection
                         pad): its extent is (H+2*a_pad) x (W+2*a_pad); output (y,x) reads (y+a_pad+dy, x+a_pad+dx) */
  int w_dup;          /* 0/1: weights span all C channels.  2: the C channels are [hi half | lo half] of an fp16
                         hi+lo split activation (see o_split) and BOTH halves use the same weights
                         [(B,) Nrows, taps*(C/2)] — the low half restores the bits fp16 rounding dropped */
} chb_conv_seg;

enum { CHB_EPI_PLAIN = 0, CHB_EPI_MODULATE = 1 };
enum { CHB_ACT_NONE = 0, CHB_ACT_RELU = 1, CHB_ACT_LRELU = 2, CHB_ACT_TANH = 3 };
enum { CHB_F16 = 0, CHB_F32 = 1 };

typedef struct {
  int B, H, W;        /* images, rows, columns of the output (== input) grid                     */
  int TW, TH, TB;     /* pixel tile: TW*TH*TB <= 128 rows of the MMA                             */
  int nseg;
  chb_conv_seg seg[4];
  int N;              /* valid output columns                                                    */
  int Nrows;          /* weight rows, multiple of BN                                             */
  int BN;             /* N tile: 16..256, multiple of 16                                         */
  int epi, act;
  const float* bias;  /* [Nrows] (or [B,Nrows] when bias_per_image), may be NULL                 */
  int bias_per_image;
  /* PLAIN: out = act(acc + bias (+ res)) */
  void* out; int out_dtype;
  int64_t o_sb, o_sy, o_sx, o_sn; /* element strides                                             */
  int o_ngroup; int64_t o_sgroup; /* n -> (n / ngroup) * sgroup + (n % ngroup) * o_sn             */
  const float* res; int64_t r_sb, r_sy, r_sx; int r_shift; /* fp32 NHWC residual read at (y>>s,x>>s) */
  /* MODULATE (ACE, normalization.py:111-112,177-187): tile columns are [gamma | beta] halves;
     out = act( (x*a + noise*nv + c) * (1+gamma) + beta ), out fp16 */
  const float* x; int64_t x_sb, x_sy, x_sx; int x_shift;
  const float* noise; /* [B, W, H] — the reference's randn(B,W,H,1) plane, read transposed; may be NULL */
  const float* chan;  /* planar [3][C]: a = rstd | c = -mean*rstd | nv = noise_var*rstd (C = N/2)         */
  /* fp16 hi+lo split output (out_dtype CHB_F16, channels-last): besides hi = fp16(v) at channel n, the rounding
     residual lo = fp16(v - hi) is stored at channel n + o_lo_off, so that a consumer segment with w_dup = 2 sees
     v to ~2^-22.  0 = off. */
  int o_split; int64_t o_lo_off;
  /* Split-K for launches with few output tiles and a long K (8x8 / 16x16 convs at small batch, the 2x2 / 4x4 shape
     layers): ksplit > 1 CTAs share one output tile, each accumulating a slice of every segment's channel chunks; the
     partial accumulators meet in ks_ws (chb_conv_ksplit_workspace_bytes(); zero its first 4096 bytes once) and the CTA
     that arrives last adds them in split order and runs the epilogue — deterministic, but a different fp32
     association than ksplit = 1.  PLAIN epilogue only; 0 / 1 = off. */
  int ksplit; void* ks_ws;
} chb_conv_desc;

enum { CHB_IMPL_TCGEN05 = 0, CHB_IMPL_SIMT_DEBUG = 1 };
/* sizeof() of the ABI structs as this library was compiled: 0 chb_conv_seg, 1 chb_conv_desc, 2 chb_gen_config,
 * 3 chb_mlp_layer (-1: unknown index).  A foreign-language binding checks its mirror against these at load time. */
int chb_struct_size(int which);
/* One-shot: encode tensor maps and launch. impl selects the tcgen05 kernel or the slow SIMT checker kernel. */
int chb_conv_run(const chb_conv_desc* d, int impl, void* stream);
/* Bytes of a split-K workspace that serves every launch with ksplit * (output tiles) <= max_ctas. */
int64_t chb_conv_ksplit_workspace_bytes(int max_ctas);

/* ------------------------------------------------------------------------------------------
 * Operator 2: label map -> one-hot pyramid (pix2pix_model.py:119-144 scatter_, and the nearest
 * F.interpolate of normalization.py:115 / generator.py:75).  Integer work, bit exact.
 * labels u8 [B,S,S]; for each level l: out[l] fp16 [B,r_l,r_l,32] with r_l = S >> shift[l].
 * ------------------------------------------------------------------------------------------ */
int chb_onehot_pyramid(const uint8_t* labels, int B, int S, int nlevels, const int* shifts, void* const* outs,
                       int nclass, void* stream);

/* Standard-normal noise planes (replaces torch.randn of normalization.py:111), Philox4x32-10 + Box-Muller. */
int chb_noise_fill(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* fp32 -> fp16 conversion of a contiguous buffer (style codes). */
int chb_f32_to_f16(const float* in, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Operator 3: fused dense chain for the colour/texture MLPs (color_texture_branch/model_eigengan.py:62-83,
 * model.py:108-127, predictor/predictor_model.py:32-41 with my_torchlib/module.py:56-64 LinearBlock).
 * Per layer:  x <- act_pre(x + [(inj_l * z[zoff:zoff+nb]) @ inj_u + inj_mu]);  y = act_post(wt^T x + bias).
 * wt is stored transposed [in_dim][out_dim]; eval-mode BatchNorm1d is folded into wt/bias by the host packer.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int in_dim, out_dim;
  const float* wt;      /* [in_dim][out_dim] */
  const float* bias;    /* [out_dim] or NULL */
  int pre_act, post_act;/* CHB_ACT_* */
  const float* inj_u;   /* [inj_nb][in_dim] or NULL  (SubspaceLayer.U,  model_eigengan.py:17) */
  const float* inj_l;   /* [inj_nb]                  (SubspaceLayer.L)  */
  const float* inj_mu;  /* [in_dim]                  (SubspaceLayer.mu) */
  int inj_nb, inj_zoff; /* number of basis vectors, offset into the per-sample z vector */
} chb_mlp_layer;
/* x fp32 [B, layers[0].in_dim], z fp32 [B, zdim] (or NULL), out fp32 [B, layers[n-1].out_dim]; device pointers. */
int chb_mlp_forward(const chb_mlp_layer* layers, int nlayers, const float* x, const float* z, int zdim, float* out,
                    int B, void* stream);

/* ------------------------------------------------------------------------------------------
 * The generator (generator.py:14-109 SPADEGenerator, 'normal' upsampling: 7 SPADE ResBlocks).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int ngf;        /* 64 in the reference (base_options.py:39)                                    */
  int label_nc;   /* 19                                                                          */
  int crop;       /* 256 (or 512); multiple of 32                                                */
  int style_len;  /* 512                                                                         */
  int max_batch;  /* workspace is sized for this many images per forward                         */
  unsigned precision; /* where fp16 operand rounding is compensated by an fp16 hi+lo split (0 = nowhere):
                         CHB_PREC_IMG      conv_img input and weights (generator.py:107-108)
                         CHB_PREC_SHORTCUT conv_s input and weights in every learned-shortcut block (architecture.py:86-93)
                         CHB_PREC_H1(i) / CHB_PREC_H0(i)  the input of conv_1 / conv_0 of block i (0 = head_0 .. 6 = up_3)
                         CHB_PREC_W(i)     the fp16 rounding residual of block i's conv_0 / conv_1 weights              */
} chb_gen_config;
enum { CHB_PREC_IMG = 1u, CHB_PREC_SHORTCUT = 2u };
#define CHB_PREC_H1(i) (1u << (8 + (i)))
#define CHB_PREC_H0(i) (1u << (16 + (i)))
#define CHB_PREC_W(i) (1u << (24 + (i)))   /* conv_0 / conv_1 WEIGHTS of block i: one more K-segment a_hi * w_lo each */

typedef struct chb_generator chb_generator;

int chb_generator_create(const chb_gen_config* cfg, chb_generator** out);
void chb_generator_destroy(chb_generator* g);

/* Packed-weight blob layout: the library owns the layout, the host packer fills it by name. */
int chb_generator_num_tensors(const chb_generator* g);
int chb_generator_tensor_info(const chb_generator* g, int i, char* name, int name_cap, int64_t* offset,
                              int64_t* nbytes, int* dtype /* CHB_F16 / CHB_F32 */);
int64_t chb_generator_blob_bytes(const chb_generator* g);
int64_t chb_generator_workspace_bytes(const chb_generator* g);
/* Both pointers are device memory that must stay alive and unchanged while the generator is used. */
int chb_generator_bind(chb_generator* g, const void* blob, void* workspace);

/* Device-resident forward.  labels u8 [B,crop,crop] (values < label_nc), codes fp32 [B,label_nc,style_len],
 * noise fp32: the 18 planes of one forward concatenated in call order (per block ace_s, ace_0, ace_1), each
 * [B, r, r] in the reference's (w,h) order, or NULL to draw them on the device from `seed`.
 * out fp32 [B,3,crop,crop] NCHW in [-1,1]. */
int chb_generator_forward(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                          uint64_t seed, float* out, int B, int impl, void* stream);
/* Same with device-drawn noise, replayed from ONE captured CUDA graph per batch size: the latency path of the
 * reference's interactive callers (HairEditor.gen_img, hair_editor.py:159-179, one image per call).  The first call
 * for a batch size captures; later calls cost one cudaGraphLaunch.  Results are identical to chb_generator_forward
 * with the same seed. */
int chb_generator_forward_graph(chb_generator* g, const uint8_t* labels, const float* codes, uint64_t seed, float* out,
                                int B, void* stream);
/* Same, host buffers in and out (pinned or pageable); copies are issued on `stream` and the call returns
 * after the output has landed in `out_host`. */
int chb_generator_forward_host(chb_generator* g, const uint8_t* labels_host, const float* codes_host,
                               const float* noise_host, uint64_t seed, float* out_host, int B, int impl,
                               void* stream);
/* Streamed form of chb_generator_forward_host for callers that render many batches (validation_in_train.py:28-33,
 * script_find_direction.py:55-74 loop gen_img): returns once the work is enqueued.  The H2D copies run on an internal
 * stream and overlap the previous batch's kernels, the D2H copy of the image overlaps the next batch's kernels; inputs
 * and output are double buffered inside the workspace.  Host buffers should be pinned and must stay valid until
 * chb_generator_host_sync() returns; noise is drawn on the device from `seed`. */
int chb_generator_forward_host_async(chb_generator* g, const uint8_t* labels_host, const float* codes_host,
                                     uint64_t seed, float* out_host, int B, void* stream);
int chb_generator_host_sync(chb_generator* g);
/* Same as chb_generator_forward (tcgen05 path) with a CUDA event recorded on `stream` between consecutive conv
 * launches: fills ms[i] / flops[i] (tensor-core FLOPs issued) for launch i and returns the number of launches
 * (or a negative error).  Synchronises on the last event.  Used by bench.py for the live roofline figure. */
int chb_generator_forward_timed(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                                uint64_t seed, float* out, int B, void* stream, float* ms, double* flops, int cap);
/* Name of conv launch i of the schedule for batch B ("up_3.ace_0.gamma_beta_mod", ...). */
int chb_generator_step_name(chb_generator* g, int B, int i, char* name, int cap);
int64_t chb_generator_noise_floats(const chb_generator* g, int B);
/* Number of kernel launches one forward issues (for bench accounting). */
int chb_generator_launches(const chb_generator* g);
/* FLOPs one forward of B images issues on the tensor cores (2*M*N*K over every tile, padding included). */
double chb_generator_flops(const chb_generator* g, int B);
/* Debug: run only the first n conv launches of the schedule (n < 0: all). */
int chb_generator_set_step_limit(chb_generator* g, int n);
/* Debug: copy an intermediate tensor by name ("x_head_0", ...) for parity tests; returns element count. */
int64_t chb_generator_debug_tensor(const chb_generator* g, const char* name, int B, void** dev_ptr, int* dtype);

/* ------------------------------------------------------------------------------------------
 * The style encoder (architecture.py:154-207 Zencoder; called through Pix2PixModel.forward(data, 'style_code'),
 * pix2pix_model.py:69-72, from HairEditor.get_code, hair_editor.py:149-157).
 * img fp32 [B,3,crop,crop] in [-1,1] (NCHW, the reference layout), labels u8 [B,crop,crop]
 * -> style codes fp32 [B,label_nc,512]; rows of absent classes are zero.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int crop;       /* 256 */
  int label_nc;   /* 19 */
  int max_batch;
} chb_zenc_config;
typedef struct chb_zencoder chb_zencoder;
int chb_zencoder_create(const chb_zenc_config* cfg, chb_zencoder** out);
void chb_zencoder_destroy(chb_zencoder* z);
int chb_zencoder_num_tensors(const chb_zencoder* z);
int chb_zencoder_tensor_info(const chb_zencoder* z, int i, char* name, int name_cap, int64_t* offset, int64_t* nbytes,
                             int* dtype);
int64_t chb_zencoder_blob_bytes(const chb_zencoder* z);
int64_t chb_zencoder_workspace_bytes(const chb_zencoder* z);
int chb_zencoder_bind(chb_zencoder* z, const void* blob, void* workspace);
int chb_zencoder_forward(chb_zencoder* z, const float* img, const uint8_t* labels, float* out, int B, void* stream);
int chb_zencoder_forward_host(chb_zencoder* z, const float* img_host, const uint8_t* labels_host, float* out_host,
                              int B, void* stream);

/* ------------------------------------------------------------------------------------------
 * Face parsing: BiSeNet on a ResNet-18 context path (external_code/face_parsing/model.py:230-274, resnet.py:21-93),
 * driven as FaceParsing.parsing_img + swap_parsing_label_to_celeba_mask + HairEditor.get_mask do it
 * (my_parsing_util.py:31-54, hair_editor.py:331-335; the reference runs it on the CPU at 512x512 for every image).
 * img u8 [B,size,size,3] RGB, already resized to the network size (the PIL bilinear resize stays on the host)
 * -> label map u8 [B,out_size,out_size]: argmax of the bilinearly x8-upsampled logits sampled at every
 * (size/out_size)-th pixel (cv2 INTER_NEAREST), mapped through the blob's 19-entry label LUT.
 * Eval BatchNorm is folded into conv weights/bias by the host packer (ctrlhair_b200/bisenet.py).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int size;       /* 512 (my_parsing_util.py:35); a multiple of 128 */
  int n_classes;  /* 19 */
  int max_batch;
} chb_bisenet_config;
typedef struct chb_bisenet chb_bisenet;
int chb_bisenet_create(const chb_bisenet_config* cfg, chb_bisenet** out);
void chb_bisenet_destroy(chb_bisenet* n);
int chb_bisenet_num_tensors(const chb_bisenet* n);
int chb_bisenet_tensor_info(const chb_bisenet* n, int i, char* name, int name_cap, int64_t* offset, int64_t* nbytes,
                            int* dtype);
int64_t chb_bisenet_blob_bytes(const chb_bisenet* n);
int64_t chb_bisenet_workspace_bytes(const chb_bisenet* n);
int chb_bisenet_launches(const chb_bisenet* n);
int chb_bisenet_bind(chb_bisenet* n, const void* blob, void* workspace);
/* Device buffers.  logits_out (optional): the 1/8-resolution logits fp32 [B,size/8,size/8,32] (19 valid channels). */
int chb_bisenet_forward(chb_bisenet* n, const uint8_t* img, uint8_t* mask, int out_size, float* logits_out, int B,
                        void* stream);
/* Pillow's Image.resize(..., BILINEAR) for uint8 [B,H,W,C] images, bit exact, on the device (my_parsing_util.py:35 resizes
 * every image to 512x512 with PIL before parsing).  tmp: scratch of B*H*OW*C bytes.  The four tables are Pillow's
 * per-output-pixel windows ([O][2] = first input index, count) and 22-bit fixed-point coefficients ([O][ks]) for the x and
 * the y pass, computed on the host in double precision (ctrlhair_b200/bisenet.py: pil_bilinear_tables). */
int chb_pil_resize_bilinear(const uint8_t* in, uint8_t* tmp, uint8_t* out, int B, int H, int W, int C, int OH, int OW,
                            const int* xbounds, const int* xcoef, int xks, const int* ybounds, const int* ycoef, int yks,
                            void* stream);
/* Host buffers in and out; returns after the label map has landed in mask_host. */
int chb_bisenet_forward_host(chb_bisenet* n, const uint8_t* img_host, uint8_t* mask_host, int out_size, int B,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * The shape branch generator (shape_branch/model.py:146-199; HairEditor.mask_generator, hair_editor.py:96;
 * call sites ui/backend.py:85-89,282-283,312,416-419).  Masks are fp32 NCHW one-hot planes as
 * shape_util.split_hair_face (shape_util.py:23-26) produces them; crop is fixed at 256 like the reference.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int crop;       /* 256 */
  int max_batch;
} chb_shape_config;
typedef struct chb_shape chb_shape;
int chb_shape_create(const chb_shape_config* cfg, chb_shape** out);
void chb_shape_destroy(chb_shape* z);
int chb_shape_num_tensors(const chb_shape* z);
int chb_shape_tensor_info(const chb_shape* z, int i, char* name, int name_cap, int64_t* offset, int64_t* nbytes,
                          int* dtype);
int64_t chb_shape_blob_bytes(const chb_shape* z);
int64_t chb_shape_workspace_bytes(const chb_shape* z);
int chb_shape_bind(chb_shape* z, const void* blob, void* workspace);
/* net 0: hair encoder, mask [B,1,256,256] -> out [B,32] = mean(16) ++ raw std head(16) (model.py:102-107; the caller
 * takes |.| of the second half); net 1: face encoder, mask [B,18,256,256] -> out [B,1024]. */
int chb_shape_encode(chb_shape* z, int net, const float* mask, float* out, int B, void* stream);
/* The same from a label map uint8 [B,S,S] (255 = no label): the one-hot planes of mask_label_to_one_hot + split_hair_face
 * (shape_branch/shape_util.py:6-26, ui/backend.py:81-84) are synthesised inside the input gather. */
int chb_shape_encode_labels(chb_shape* z, int net, const uint8_t* labels, float* out, int B, void* stream);
/* forward_decode_by_code (model.py:195-199): -> softmax mask fp32 [B,19,256,256]. */
int chb_shape_decode(chb_shape* z, const float* hair_code, const float* face_code, float* mask_out, int B,
                     void* stream);
/* forward_decode_by_code followed by mask_one_hot_to_label (ui/backend.py:89-90,312-313; shape_util.py:17-20) in one
 * call: uint8 labels [B,S,S] = argmax of the softmax probabilities, without writing the [B,19,S,S] tensor. */
int chb_shape_decode_labels(chb_shape* z, const float* hair_code, const float* face_code, uint8_t* labels_out, int B,
                            void* stream);
/* forward_hair_decoder (net 0, model.py:175-178; input cat([face_code, hair_code])) and forward_face_decoder (net 1,
 * model.py:180-182; hair_code ignored, may be NULL): logits fp32 NCHW, [B,1,256,256] / [B,18,256,256]
 * (ui/backend.py:416 directly_change_hair_mask reads the face logits). */
int chb_shape_decode_logits(chb_shape* z, int net, const float* hair_code, const float* face_code, float* logits_out,
                            int B, void* stream);
/* forward_decoder (model.py:184-187; ui/backend.py:419): softmax over [face[:13], hair, face[13:]] of caller-provided
 * logits (fp32 NCHW) -> mask fp32 [B,19,256,256]. */
int chb_shape_softmax(chb_shape* z, const float* hair_logit, const float* face_logit, float* mask_out, int B,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Colour/texture training step, config 045 (color_texture_branch/train.py:115-148 loop body;
 * solver.py:85-117 forward, :218-245 forward_d incl. the WGAN-GP double backward :204-216, :119-166 forward_g;
 * my_torchlib/train_utils.py:54-89 train(); Adam solver.py:52-55).  fp32 throughout.
 *
 * One call = forward + losses + backward of ONE sub-step (which = 0: discriminator update, 1: generator update)
 * on this rank's batch; gradients land in the `grads` region of the state buffer so that the host can all-reduce
 * them (DDP, solver.py:68-74) before chb_cttrain_adam applies the update.  Every random draw of the reference
 * (three shuffles + encoder-noise coin of solver.py:98-111, alpha_gp of :199) is an explicit input.
 * The whole sub-step is captured once into a CUDA graph and replayed.
 * ------------------------------------------------------------------------------------------ */
enum { CHB_CTT_D = 0, CHB_CTT_G = 1, CHB_CTT_FROZEN = 2 };
enum {  /* slots of the losses[] output (unweighted values, the reference's loss_dict keys) */
  CHB_CTT_L_ADV = 0, CHB_CTT_L_GP, CHB_CTT_L_INFO, CHB_CTT_L_REC, CHB_CTT_L_MOMENT_1, CHB_CTT_L_MOMENT_2,
  CHB_CTT_L_INFO_CURLINESS, CHB_CTT_L_RGB, CHB_CTT_L_PCA_STD, CHB_CTT_L_CLS_CURLINESS, CHB_CTT_L_ORTHOGONAL,
  CHB_CTT_L_TOTAL, CHB_CTT_NUM_LOSSES
};
typedef struct {
  int batch;            /* samples per call on this rank (cfg.batch_size = total_batch_size / gpu_num)        */
  float lambda_adv, lambda_gp, lambda_info, lambda_info_curliness, lambda_rec, lambda_rgb, lambda_pca_std,
      lambda_moment_1, lambda_moment_2, lambda_cls_curliness, lambda_orthogonal; /* config.py:16-39,52-96   */
  float lr, beta1, beta2, eps;                                                      /* 2e-4, .5, .999, 1e-8    */
  int use_graph;        /* 0: plain launches; 1: each sub-step as an explicit CUDA graph built on first use, one node per
                           operation and an edge per real data dependency (the default of the Python host); 2: each
                           sub-step as ONE persistent cooperative kernel walking the operation list with grid barriers
                           between dependent operations only.  Same arithmetic in all three.                       */
} chb_cttrain_config;
typedef struct {  /* device pointers, fp32 unless noted; rows = batch */
  const float* code;             /* [B,512]  data['code']                                   */
  const float* rgb_mean;         /* [B,3]                                                   */
  const float* pca_std;          /* [B,1]                                                   */
  const float* noise;            /* [B,8]    generate_noise (train_utils.py:44-51)          */
  const float* noise_curliness;  /* [B,1]    |N(0,1)| * label                               */
  const float* curliness_label;  /* [B,1]    -1 / +1 as float                               */
  const int* perm_rgb;           /* [B] int32: first shuffle  (rgb_mean, pca_std)           */
  const int* perm_curliness;     /* [B] second shuffle (noise_curliness, curliness_label)   */
  const int* perm_noise;         /* [B] third shuffle  (noise)                              */
  const float* alpha_gp;         /* [B,1] U(0,1); D sub-step only (may be NULL for G)       */
  int noise_from_encoder;        /* the gan_input_from_encoder_prob coin (solver.py:107-111) */
} chb_cttrain_batch;
typedef struct chb_cttrain chb_cttrain;
int chb_cttrain_create(const chb_cttrain_config* cfg, chb_cttrain** out);
void chb_cttrain_destroy(chb_cttrain* t);
/* Parameter table: names are "D." / "G." + the reference state_dict key (trainable), "P." / "C." + key for the frozen
 * rgb / curliness predictors (fc weights with the eval BatchNorm1d folded in by the host packer).  `offset` and
 * `numel` are in floats inside the region of `group` (CHB_CTT_D / _G / _FROZEN). */
int chb_cttrain_num_tensors(const chb_cttrain* t);
int chb_cttrain_tensor_info(const chb_cttrain* t, int i, char* name, int name_cap, int64_t* offset, int64_t* numel,
                            int* group);
/* State buffer (fp32, caller-owned device memory): [params D | params G | grads D | grads G | adam m | adam v |
 * frozen].  region: 0 params, 1 grads, 2 adam m, 3 adam v (group D or G), or group FROZEN (region ignored). */
int64_t chb_cttrain_state_floats(const chb_cttrain* t);
int chb_cttrain_region(const chb_cttrain* t, int region, int group, int64_t* offset, int64_t* numel);
int64_t chb_cttrain_workspace_bytes(const chb_cttrain* t);
int chb_cttrain_bind(chb_cttrain* t, float* state, void* workspace);
/* forward + losses + backward of one sub-step; losses_out: device float[CHB_CTT_NUM_LOSSES] (slots a sub-step does
 * not compute are written as 0).  Gradients of the stepped net replace the previous content of its grads region. */
int chb_cttrain_step(chb_cttrain* t, int which, const chb_cttrain_batch* batch, float* losses_out, void* stream);
/* Adam update of one net from its grads region (after the optional all-reduce); advances that net's step count. */
int chb_cttrain_adam(chb_cttrain* t, int which, void* stream);
int chb_cttrain_launches(const chb_cttrain* t, int which);
/* After the first use of a sub-step: its number of operations and, for use_graph 1, of dependency edges in its graph,
 * for use_graph 2, of grid barriers in its persistent kernel. */
int chb_cttrain_schedule(const chb_cttrain* t, int which, int* n_ops, int* n_barriers);

/* ------------------------------------------------------------------------------------------
 * Operator 8: post-processing either side of the generator.  Device pointers; integer
 * results are bit exact, the Poisson solve is fp64 conjugate gradients to a relative residual `tol`.
 * ------------------------------------------------------------------------------------------ */
/* hair_editor.py:273-288: generator output float [B,3,H,W] in [-1,1] -> uint8 [B,H,W,3] = (x*127.5+127.5).astype(uint8). */
int chb_image_to_u8(const float* img, uint8_t* out, int B, int H, int W, void* stream);
/* hair_editor.py:297-306: res_mask = (target_parsing == 13) | (face_parsing == 13), dilated by cv2's 13x13 ellipse
 * (5x5 where target_parsing is background).  Parsings uint8 [B,H,W].  Writes res_mask_dilated (0/1) and/or
 * solve_mask = 1 - res_mask_dilated (what poisson_blending receives); either may be NULL. */
int chb_blend_mask(const uint8_t* target_parsing, const uint8_t* face_parsing, uint8_t* res_mask_dilated,
                   uint8_t* solve_mask, int B, int H, int W, void* stream);
/* poisson_blending.py:29-87 (replaces the lil_matrix assembly + three spsolve calls).  source, target, out uint8
 * [B,H,W,3]; mask uint8 [B,H,W], non-zero = keep the source's gradients.  ceil(H/8)*W <= 8192, W <= 512.
 * stats (optional) float [B*3][2] = CG iterations, final relative residual per (image, channel).
 * lut_fwd (optional, device, double[256]) = v**(1/2.2) and lut_known (optional, device, uint8[256]) =
 * uint8((v**(1/2.2))**2.2) as the CALLER's host computes them: pow() differs by an ulp between libm / SVML builds of
 * numpy, which decides whether an untouched pixel v comes back as v or v-1; NULL = the device's pow. */
int chb_poisson_blend(const uint8_t* source, const uint8_t* target, const uint8_t* mask, uint8_t* out, int B, int H,
                      int W, int with_gamma, double tol, int max_iter, float* stats, const double* lut_fwd,
                      const uint8_t* lut_known, void* stream);
/* The coarse level of chb_poisson_blend's preconditioner on its own (a test hook; the solve calls it internally for
 * 256-column images): per image, the inverse of the Galerkin operator P^T A P of poisson_blending.py:44-70's matrix on
 * 16 x 16-pixel aggregates.  mask uint8 [B,H,256]; inv_f16 IEEE half [B][256][256], row a, entry for aggregate a' at
 * column (a' & 15) * 16 + (a' >> 4), scaled by 256. */
int chb_poisson_coarse_inverse(const uint8_t* mask, uint16_t* inv_f16, int B, int H, void* stream);
/* hair_editor.py:257-308 HairEditor.postprocess_blending in one call: face_img uint8 [B,H,W,3], res_img float
 * [B,3,H,W] (generator output), parsings uint8 [B,H,W] -> out uint8 [B,H,W,3] (+ res_mask_dilated [B,H,W], optional).
 * blending == 0: out = uint8 image of res_img only.  workspace: chb_postprocess_workspace_bytes(B,H,W) device bytes. */
int64_t chb_postprocess_workspace_bytes(int B, int H, int W);
int chb_postprocess_blending(const uint8_t* face_img, const float* res_img, const uint8_t* face_parsing,
                             const uint8_t* target_parsing, uint8_t* out, uint8_t* res_mask_dilated, void* workspace,
                             int B, int H, int W, int blending, double tol, int max_iter, float* stats,
                             const double* lut_fwd, const uint8_t* lut_known, void* stream);
/* ui/backend.py:98-101,117-125 (cv2.cvtColor(c.astype('uint8'), COLOR_RGB2HSV)) and :108-115 (COLOR_HSV2RGB) on n
 * colour triples: OpenCV's 8-bit algorithms (H in 0..179), bit exact.  Pass exactly one of rgb_f32 / rgb_u8. */
int chb_rgb_to_hsv(const float* rgb_f32, const uint8_t* rgb_u8, uint8_t* hsv, int64_t n, void* stream);
int chb_hsv_to_rgb(const uint8_t* hsv, uint8_t* rgb, int64_t n, void* stream);
/* shape_branch/shape_util.py:17-20 mask_one_hot_to_label: float [B,C,hw] -> uint8 [B,hw] (first maximum, 255 if all 0)
 * and :6-14 mask_label_to_one_hot: uint8 [B,hw] (255 = none) -> float [B,C,hw]. */
int chb_onehot_to_label(const float* one_hot, uint8_t* labels, int B, int C, int64_t hw, void* stream);
int chb_label_to_onehot(const uint8_t* labels, float* one_hot, int B, int C, int64_t hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif
