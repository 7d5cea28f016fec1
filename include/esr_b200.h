/* esr_b200 — C ABI of the B200-native x4 efficient-SR inference engine.
 *
 * This is the drop-in boundary for the hot path of ofsoundof/NTIRE2022_ESR: the call
 * `model(img_lq)` made by `forward()` (reference test_demo.py:364-367) on the module returned by
 * `select_model()` (test_demo.py:13-341).  The reference has no FFI of its own (it is pure PyTorch);
 * the entry points below are what a Python/ctypes binding of that path needs, one per step of
 * `select_model` (construct -> load_state_dict(strict=True) -> eval/to(device)) and of `forward`.
 * The Python mirror lives in ntire2022_esr_b200/engine.py; INTEGRATION.md shows the stub a
 * maintainer of the reference would add to test_demo.py.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * ESR_E_* code and never throws/aborts; esr_last_error() gives the message.  A handle is bound to
 * one CUDA device, calls on it are stream-ordered and not re-entrant.  Device buffers (input,
 * output, workspace) are owned by the caller; packed weights and TMA descriptors by the engine.
 */
#ifndef ESR_B200_H_
#define ESR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct esr_engine esr_handle;

/* architectures: the four networks SURVEY.md section 8(a) puts on the hot path, plus the pruned RFDN of row N1
 * (models/team40_rfdn_pruned.py: RFDBs without the inner residual adds, ESA width 12) */
enum { ESR_ARCH_IMDN = 0, ESR_ARCH_RFDN = 1, ESR_ARCH_RLFN = 2, ESR_ARCH_BSRN = 3, ESR_ARCH_RFDN_PRUNED = 4, ESR_ARCH_FMEN = 5 };
/* I/O + storage dtype of a forward call (math is fp32-accumulate in both) */
enum { ESR_DTYPE_F32 = 0, ESR_DTYPE_F16 = 1 };
enum {
  ESR_OK = 0,
  ESR_E_INVALID = -1,   /* bad argument / unsupported shape */
  ESR_E_STATE = -2,     /* call order (e.g. forward before finalize) */
  ESR_E_WEIGHTS = -3,   /* missing / unexpected / mis-shaped state-dict entry (strict=True) */
  ESR_E_CUDA = -4,      /* CUDA runtime or driver error */
  ESR_E_NOGPU = -5      /* no sm_100 device: the engine has no CPU fallback */
};

/* Replaces the module constructors at test_demo.py:22 (IMDN(nc=64, nb=8)), :29 (RFDN()),
 * :57 (RLFN_cut()), :155 (BSRN(num_feat=48, num_block=5)), :180 (RFDN40() = ESR_ARCH_RFDN with nf = 40) and
 * :307 (RFDNPrune(nf=40)).  nf / nblocks <= 0 pick those defaults.  `device` is the CUDA ordinal the handle is bound to. */
int esr_create(esr_handle** out, int arch, int nf, int nblocks, int device);

/* Replaces `model.load_state_dict(torch.load(path), strict=True)` (test_demo.py:23,30,58,157):
 * called once per state-dict entry with its exact reference name, a HOST fp32
 * pointer and its shape.  Unknown names / wrong shapes are reported by esr_finalize. */
int esr_load_weights(esr_handle* h, const char* name, const float* host_ptr, const int64_t* shape, int ndim);

/* Replaces `.eval()` + `.to(device)` (test_demo.py:336-340): checks the entry set strictly, packs
 * the weights into the engine's layouts (fp32 tables for the CUDA-core kernels, pre-swizzled fp16
 * K-major blocks for the tcgen05 kernels) and uploads them. */
int esr_finalize(esr_handle* h);

/* Scratch bytes esr_forward needs for a (B,3,H,W) input of `dtype`; 0 on error.  The workspace holds state between
 * calls (cleared pad lanes): the caller must not write to it; a call with another shape / dtype re-clears it. */
size_t esr_workspace_bytes(esr_handle* h, int B, int H, int W, int dtype);

/* Replaces `model(img_lq)` (test_demo.py:367, and :380 for each tile): in = (B,3,H,W) NCHW, values
 * in [0,data_range]; out = (B,3,4H,4W) NCHW; both DEVICE pointers of `dtype`.  Asynchronous on
 * `stream` (a cudaStream_t passed as void*; NULL = default stream). */
int esr_forward(esr_handle* h, const void* in_nchw, void* out_nchw, int B, int H, int W, int dtype,
                void* workspace, size_t workspace_bytes, void* stream);

/* uint8 image in, uint8 image out: replaces the three calls the reference's run() makes per image,
 *   img_lr = util.uint2tensor4(img_lr, data_range)   (test_demo.py:423, utils/utils_image.py:190-193)
 *   img_sr = forward(img_lr, model, tile)            (test_demo.py:430)
 *   img_sr = util.tensor2uint(img_sr, data_range)    (test_demo.py:434, utils/utils_image.py:204-208)
 * in = (B,H,W,3) interleaved uint8, out = (B,4H,4W,3) interleaved uint8, both DEVICE pointers.  The scaling
 * u8 / (255 / data_range) is folded into the first convolution's load; the clamp / * 255 / data_range /
 * round-half-even / uint8 cast runs on the device on the network's output, so only 1/2 (fp16) or 1/4 (fp32) of
 * the output bytes ever leave the GPU.  `dtype` selects the engine precision (ESR_DTYPE_F16: result identical to
 * tensor2uint(model(uint2tensor4(img).half()))).  The workspace needs esr_workspace_bytes_u8() bytes. */
size_t esr_workspace_bytes_u8(esr_handle* h, int B, int H, int W, int dtype);
int esr_forward_u8(esr_handle* h, const uint8_t* in_hwc, uint8_t* out_hwc, int B, int H, int W, float data_range,
                   int dtype, void* workspace, size_t workspace_bytes, void* stream);
/* the same with HOST buffers (see esr_forward_host / esr_forward_host_async below) */
int esr_forward_host_u8(esr_handle* h, const uint8_t* in_host_hwc, uint8_t* out_host_hwc, int B, int H, int W,
                        float data_range, int dtype);
int esr_forward_host_u8_async(esr_handle* h, const uint8_t* in_host_hwc, uint8_t* out_host_hwc, int B, int H, int W,
                              float data_range, int dtype, long long* ticket);

/* Same computation with HOST buffers (pageable or pinned): the engine stages them through its own
 * device buffers on `stream` and synchronises before returning.  This is the path a non-PyTorch
 * caller binds; bench.py times it as the end-to-end number. */
int esr_forward_host(esr_handle* h, const void* in_host, void* out_host, int B, int H, int W, int dtype);

/* Pipelined form of the same path for serving loops (run() of the reference, test_demo.py:416-434, one
 * image after another): returns once the request is queued on the engine's streams; up to 4 requests are
 * in flight, each on its own compute stream and workspace, so the H2D copy, the forward and the D2H copy of
 * consecutive requests overlap and the forwards of independent requests run concurrently.  Host buffers
 * should be pinned.  *ticket identifies the request; esr_host_wait(h, ticket) blocks until its output (and
 * every earlier one) is in out_host; ticket < 0 waits for all. */
int esr_forward_host_async(esr_handle* h, const void* in_host, void* out_host, int B, int H, int W, int dtype,
                           long long* ticket);
int esr_host_wait(esr_handle* h, long long ticket);

/* Number of kernels one esr_forward of this shape launches (0 on error); used by bench.py. */
int esr_launch_count(esr_handle* h, int B, int H, int W, int dtype);
/* Name of the i-th launch of that plan ("conv_tc", "conv_generic", ...), or NULL. */
const char* esr_launch_name(esr_handle* h, int B, int H, int W, int dtype, int i);

/* Algorithmic FLOPs (2 x un-padded MACs) of the i-th launch of that plan. */
double esr_launch_flops(esr_handle* h, int B, int H, int W, int dtype, int i);

/* Measurement aid for bench.py's roofline: runs the plan of one forward launch by launch, each launch
 * repeated `reps` times (captured into a CUDA graph, so the CPU launch rate does not mask short kernels)
 * between two CUDA events on `stream`, and writes the mean milliseconds of
 * every launch to ms_out[0..n).  Returns the number of launches (>0) or a negative ESR_E_* code.
 * Same arguments as esr_forward; synchronises the stream. */
int esr_profile_launches(esr_handle* h, const void* in_nchw, void* out_nchw, int B, int H, int W, int dtype,
                         void* workspace, size_t workspace_bytes, int reps, float* ms_out, int n, void* stream);

/* Runtime knobs (tuning / A-B measurements): "tc_enable" (fp16: tcgen05 convolutions on/off),
 * "chain_enable" (0 per-layer kernels only, 1 fused 3x3 chains where the bands of one image fit the device, 2 always),
 * "chain_pw" (c5 / conv1x1 as the chain's last stage: measured slower, off), "chain_store_all", "chain_mask",
 * "use_graph", "use_pdl" (programmatic dependent launch: 1 every kernel, 2 only the small ESA kernels; measured slower),
 * "esa_front_old" (the ESA kernels before their specialisation on the ESA width), "tc_rows_per_item", "tc_acc_slots",
 * "tc_timeline", "tc_dbg_flags" (the last two select the instrumented kernel instantiations).  Changing an option drops
 * the cached launch plans. */
int esr_set_option(esr_handle* h, const char* key, int value);

/* Debug aid: with option "tc_timeline" = 1 every tcgen05 launch records clock64 stamps of block 0
 * (128 per launch: 4 roles x 32 events); copies the first n_launches records to `out`. */
int esr_debug_timeline(esr_handle* h, long long* out, int n_launches);

/* Test aid (host only, works on a handle without a GPU): the packed form of the index-th tcgen05 layer of the
 * fp16 plan - what conv_tc_kernel is handed - so that the weight packing, the algebraic folds and the MMA entry
 * list can be replayed on the CPU (tests/test_tc_packing_cpu.py).  index < 0 returns the number of layers.
 * meta[12] = {nchunks, halo, acc_cols, n_entries, ngroups, blob_bytes, 0, has_bias9, chunk_c0[4]};
 * entries[n][8] = {dy, dx, chunk, k16_steps, n, first_column, overwrite, blob_offset};
 * groups[g][8] = {col0, ncols, act, has_residual, residual_after_act, mode, slope (float bits), has_bias9};
 * bias[3][64]; bias9[9][64] (border classes, first group); blob = pre-swizzled fp16 [n x 64] K-major blocks. */
int esr_debug_tc_layer(esr_handle* h, int index, char* name, int name_cap, int32_t* meta, int32_t* entries, int32_t* groups,
                       float* bias, float* bias9, uint8_t* blob, size_t blob_cap);

/* Test aid (host only): the packed form of the index-th fused chain of the fp16 plan - what conv_chain_kernel is
 * handed (tests/test_chain_packing_cpu.py replays it in the kernel's row-stationary order).  index < 0 returns the
 * number of chains.  meta[4] = {n_layers, blob_bytes, first tcgen05 layer (esr_debug_tc_layer index), has pointwise stage};
 * layers[l][12] = {tcgen05 layer index, np, k16_steps, centre_cols, part_bytes, blob_offset, identity_tap, n0, n1,
 * group1_is_centre_block, group1_first_column, 0}.  Blob layout per layer: three dy parts (dy = +1, 0, -1), each
 * [np/8 atoms][3 dx][8 rows x 128 B, SWIZZLE_128B]; the centre block [centre_cols x 64] K-major SWIZZLE_128B (4 KB);
 * 128 bias floats (group 0 at [0, 64), group 1 at [64, 128)). */
int esr_debug_chain(esr_handle* h, int index, int32_t* meta, int32_t* layers, uint8_t* blob, size_t blob_cap);

const char* esr_last_error(esr_handle* h);
void esr_destroy(esr_handle* h);

/* Library-level queries (no handle): version string, and whether a usable sm_100 GPU is present. */
const char* esr_version(void);
int esr_device_ok(int device);

#ifdef __cplusplus
}
#endif
#endif /* ESR_B200_H_ */
