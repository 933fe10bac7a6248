/*
 * kraken_b200.h - C ABI of the B200-native line-recognition / segmentation-forward engine.
 *
 * This is the drop-in boundary for ONE hot path of mittagessen/kraken (paths relative to the
 * reference checkout):
 *
 *   rpred :  batched line images -> VGSL conv stack -> BiLSTM -> linear -> softmax -> CTC greedy decode
 *            kraken/lib/vgsl/rpred.py:210-229 (_rec_predict), kraken/lib/models.py:93-136
 *   blla  :  page -> VGSL net -> nearest upsample -> sigmoid
 *            kraken/lib/vgsl/spred.py:268-272, kraken/blla.py:121-125
 *
 * The reference has no FFI (it is pure Python over ATen); the narrowest operator boundary it has is
 * the callable `TorchVGSLModel.nn(x[N,C,H,W], seq_lens[N]) -> (y[N,C',H',W'], seq_lens')`
 * (kraken/lib/vgsl/layers.py:44-53, kraken/lib/vgsl/model.py:488-489) plus the decoder hook
 * `decoder(probs[N,C,W], seq_lens) -> [[(label,start,end,conf)]]` (kraken/lib/ctc_decoder.py:35-72).
 * Every entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * ctypes stub a kraken maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All tensors are fp32, NCHW, dense, as in the reference
 *     (kb_recognize_u8 additionally takes the uint8 images the reference's transforms start from).
 *   - arithmetic: fp32 on CUDA cores for stencil / scan / gate math; the tensor-core layers (convolutions, LSTM
 *     projections and recurrences, linear) read every fp32 operand as two fp16 planes (22 significand bits) and
 *     accumulate in fp32: logits agree with the fp32 reference to ~2e-6 relative, CTC label sequences bit for bit.
 *   - `*_on_device` flags say whether a data pointer is a CUDA device pointer (of the model's device) or
 *     a host pointer (pinned or pageable).  Host pointers are copied inside the call.
 *   - every function returns KB_OK or an error code; kb_last_error() returns a thread-local message.
 *     No exceptions cross the ABI.  Error codes map onto the reference's Python exceptions:
 *       KB_ERR_SPEC        -> ValueError            (model.py:160,187,194,230,239,796-804,870,898)
 *       KB_ERR_SHAPE       -> KrakenInputException  (models.py:113-114) / the layer's own Exception
 *       KB_ERR_ARG/STATE   -> ValueError / RuntimeError at the Python surface
 *       KB_ERR_CUDA        -> RuntimeError - there is NO CPU fallback: without a usable GPU every
 *                             compute entry point fails with this code.
 *   - a model handle is internally serialised (one call at a time per handle); different handles are
 *     independent and may be driven from different threads/streams.  The engine owns weights and
 *     workspaces; the caller owns every buffer it passes in.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Every compute call
 *     synchronises that stream before returning (results are complete; the operand-range check of the
 *     tensor-core layers, see kb_range_fallback_count, needs the stream idle).
 */
#ifndef KRAKEN_B200_H
#define KRAKEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_ABI_VERSION 4

enum {
    KB_OK = 0,
    KB_ERR_SPEC = 1,
    KB_ERR_ARG = 2,
    KB_ERR_CUDA = 3,
    KB_ERR_UNSUPPORTED = 4,
    KB_ERR_STATE = 5,
    KB_ERR_SHAPE = 6
};

/* layer kinds reported by kb_model_layer_info (leaf layers only, in execution order) */
enum {
    KB_LAYER_CONV = 1, KB_LAYER_MAXPOOL = 2, KB_LAYER_RESHAPE = 3, KB_LAYER_LSTM = 4, KB_LAYER_DROPOUT = 5,
    KB_LAYER_GROUPNORM = 6, KB_LAYER_LINEAR = 7, KB_LAYER_ADDITION = 8, KB_LAYER_IDENTITY = 9
};

typedef struct kb_model kb_model;

typedef struct kb_layer_info {
    int32_t kind;            /* KB_LAYER_*                                                        */
    int32_t out_shape[4];    /* static (batch, C, H, W) as the reference's get_shape(); 0 = variable */
    char    name[64];        /* e.g. "C_0"                                                         */
    char    path[256];       /* state-dict prefix below "nn.", e.g. "C_2 C_3 I_4.C_2 C_3.C_2"      */
    char    block[128];      /* named block, e.g. "Cr{C_0}3,3,32"                                  */
} kb_layer_info;

/* ---- library ------------------------------------------------------------------------------------ */
int         kb_abi_version(void);
const char *kb_last_error(void);
/* hash of the sources + compiler flags this library was built from (see __graft_entry__.py::source_hash); lets the tests and
 * build() notice a stale binary next to newer sources */
const char *kb_source_hash(void);
/* number of visible CUDA devices (0 when there is no driver/GPU; never fails) */
int         kb_device_count(void);

/* ---- spec -> graph (host only, usable without a GPU) ---------------------------------------------
 * replaces TorchVGSLModel.__init__/_parse/build_* (kraken/lib/vgsl/model.py:109-243, 570-902).        */
int  kb_model_create(const char *vgsl_spec, kb_model **out);
void kb_model_destroy(kb_model *m);
/* named spec "[1,48,0,1 Cr{C_0}3,3,32 ...]" == user_metadata['vgsl'] (model.py:198-199).
 * Returns the length needed (excluding NUL) or a negative error; writes at most cap bytes.          */
int  kb_model_named_spec(const kb_model *m, char *buf, size_t cap);
int  kb_model_input_shape(const kb_model *m, int32_t shape[4]);    /* (batch, channels, height, width) model.py:196 */
int  kb_model_output_shape(const kb_model *m, int32_t shape[4]);   /* TorchVGSLModel.output                         */
int  kb_model_num_layers(const kb_model *m);
int  kb_model_layer_info(const kb_model *m, int index, kb_layer_info *info);
/* parameters in state_dict() order; names are the reference's keys ("nn.C_0.co.weight", ...) */
int  kb_model_num_tensors(const kb_model *m);
int  kb_model_tensor_info(const kb_model *m, int index, char *name, size_t cap, int64_t shape[4], int32_t *ndim);
/* runtime shape of nn(x) for an input of n x C x h x w, and the seq_len arithmetic of every layer
 * (layers.py:334,387,858-859) applied to `widths`; pure integer/host work.                          */
int  kb_model_infer_dims(const kb_model *m, int32_t n, int32_t h, int32_t w, int32_t out_nchw[4]);
int  kb_model_infer_lens(const kb_model *m, int32_t n, int32_t h, int32_t w, const int32_t *widths, int32_t *out_lens);

/* ---- weights ------------------------------------------------------------------------------------
 * replaces load_state_dict(); data is fp32 host memory, copied.                                     */
int  kb_model_load_tensor(kb_model *m, const char *name, const float *data, const int64_t *shape, int32_t ndim);
/* uploads + repacks weights (HWIO conv filters, gate-interleaved LSTM matrices, folded biases) on
 * `device`.  Needs a GPU.  Must be called again after load_tensor().  replaces nn.to(device)
 * (kraken/lib/models.py:84-91, model.py:518-523).                                                  */
int  kb_model_finalize(kb_model *m, int device);
int  kb_model_device(const kb_model *m);            /* -1 before finalize */

/* ---- nn(x, seq_lens) -----------------------------------------------------------------------------
 * replaces `self.nn(line, lens)` (layers.py:44-53; callers rpred.py:225, models.py:112, spred.py:268,
 * blla.py:121).  x: n x C x h x w.  widths: n int32 or NULL (seq_lens=None).  out: the NCHW result
 * with the dims kb_model_infer_dims() reports; out_lens (n, may be NULL) receives seq_lens'.         */
int  kb_forward(kb_model *m, const float *x, int x_on_device, int32_t n, int32_t h, int32_t w,
                const int32_t *widths, float *out, int out_on_device, int32_t *out_lens, void *stream);

/* ---- fused recognition --------------------------------------------------------------------------
 * replaces _rec_predict up to (not including) codec.decode: nn -> (logits/T).softmax(1) -> greedy
 * decoder (rpred.py:225-228, models.py:112-136, ctc_decoder.py:55-72).
 * Outputs (host memory, caller allocated; max_out = capacity per line, use the output width T):
 *   labels/starts/ends [n*max_out] int32, confs [n*max_out] float, counts [n], out_lens [n].
 *   probs (optional, may be NULL): (n, C, T) like `self.outputs` (rpred.py:227); probs_on_device
 *   selects where it lives.  Fails with KB_ERR_SHAPE if the net's output height is not 1.            */
int  kb_recognize(kb_model *m, const float *lines, int lines_on_device, int32_t n, int32_t h, int32_t w,
                  const int32_t *widths, float temperature,
                  int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts,
                  int32_t max_out, int32_t *out_lens, float *probs, int probs_on_device, void *stream);

/* ---- fused recognition on uint8 line images --------------------------------
 * Same as kb_recognize, but the lines arrive as the uint8 tensors `v2.PILToTensor()` produces and the rest of
 * ImageInputTransforms runs on the device, bit-identical to the reference:
 *   v2.ToDtype(float32, scale=True)  x.to(float32).mul_(1.0/255)   kraken/lib/dataset/utils.py:148-149
 *   tensor_invert                    im.max() - im                 kraken/lib/functional_im_transforms.py:58-59
 *   zero right-padding to the batch width                           kraken/lib/vgsl/rpred.py:129-131
 * lines: n x C x h x w uint8 (host or device).  widths (n, or NULL): columns >= widths[i] are fed as 0 whatever the
 * buffer holds.  invert_max (n, or NULL = no inversion): per line the maximum pixel value of the un-padded crop
 * (im.max() * 255; 255 for anything that contains white), or a negative value to skip the inversion of that line.
 * A quarter of the host-to-device bytes of the float call.                                                       */
int  kb_recognize_u8(kb_model *m, const uint8_t *lines, int lines_on_device, int32_t n, int32_t h, int32_t w,
                     const int32_t *widths, const int16_t *invert_max, float temperature, int32_t *labels, int32_t *starts,
                     int32_t *ends, float *confs, int32_t *counts, int32_t max_out, int32_t *out_lens, float *probs,
                     int probs_on_device, void *stream);

#define KB_DTYPE_F32 0
#define KB_DTYPE_U8  1

/* ---- record assembly on the device -------------------------------------------------------------------------
 * kb_model_set_codec: label -> Unicode code point table of a 1:1 codec (`PytorchCodec.l2c_single` with one code point per label,
 *   kraken/lib/codec.py:164-172); l2c[label] = 0 for labels outside the codec.  n_labels = 0 clears it.
 * kb_recognize_records: kb_recognize (dtype KB_DTYPE_F32) / kb_recognize_u8 (KB_DTYPE_U8; invert_max as there) whose CTC collapse writes
 *   records instead of raw labels: codepoints[i][j] = l2c[label] (0: not in the codec - `decode` skips those unless strict), and
 *   starts / ends are positions in the ORIGINAL line image exactly as `_scale_val` yields them (kraken/lib/vgsl/rpred.py:138-157,231):
 *       net_scale = widths[i] / out_lens[i];   in_scale = orig_widths[i] / (widths[i] - 2 * padding)
 *       pos = int(round(min(max((t * net_scale - padding) * in_scale, 0), orig_widths[i] - 1)))        (Python doubles, round-half-even)
 *   orig_widths (n): width of each line image before ImageInputTransforms; padding: the transform's horizontal padding.
 *   confs / counts / out_lens as kb_recognize.  What is left for the host is ''.join(map(chr, codepoints)).                        */
int  kb_model_set_codec(kb_model *m, const uint32_t *l2c, int32_t n_labels);
int  kb_recognize_records(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w,
                          const int32_t *widths, const int16_t *invert_max, float temperature, const int32_t *orig_widths, int32_t padding,
                          uint32_t *codepoints, int32_t *starts, int32_t *ends, float *confs, int32_t *counts, int32_t max_out,
                          int32_t *out_lens, void *stream);

/* ---- forced alignment of known transcriptions (the `return_logits` consumer) -------------------------------------
 * Replaces, per batch of lines, the numeric part of `ForcedAlignmentTaskModel.predict` (kraken/tasks/align.py:104-137):
 *     emission = record.logits.squeeze().log_softmax(0).T      align.py:119  (record.logits = the (C, T) softmax output of
 *                                                              kraken/lib/vgsl/rpred.py:226-227,200 - the second softmax is the reference's)
 *     get_trellis (align.py:170-191), backtrack (:194-229), merge_repeats (:232-249), `_scale_val` of the borders (:128-132)
 * Everything stays on the device between the network and the segments; the probabilities are never copied to the host.
 * lines / dtype / widths / invert_max / temperature: as kb_recognize_records.
 * tokens: the label sequences `codec.encode(text)` of all lines, concatenated; tok_off (n + 1): tok_off[i] .. tok_off[i + 1] are line
 *   i's labels (tok_off[0] = 0).  A line without labels is an error (KB_ERR_ARG; the reference raises IndexError on it).
 * orig_widths (n, or NULL) / padding: NULL = segment borders in output frames [start, end); otherwise `_scale_val(border, 0, width)`
 *   positions in the original line image, with net_scale / in_scale as align.py:128 and rpred.py:185-187 set them.
 * Outputs (host, caller allocated, max_seg >= the longest label sequence): seg_token [n * max_seg] = index INTO THE LINE'S LABEL SEQUENCE
 *   (merge_repeats labels a segment ground_truth[token_index]), seg_start / seg_end, seg_score = mean frame probability of the run;
 *   seg_counts [n] = number of segments, or -1: fewer output frames than 2 * len(labels) - the reference emits an empty record
 *   (align.py:113-117), or -2: the backtrack ran out of frames - ValueError('Failed to align') (align.py:228).  out_lens (n, optional).   */
int  kb_forced_align(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w,
                     const int32_t *widths, const int16_t *invert_max, float temperature, const int32_t *tokens, const int32_t *tok_off,
                     const int32_t *orig_widths, int32_t padding, int32_t *seg_token, int32_t *seg_start, int32_t *seg_end,
                     float *seg_score, int32_t *seg_counts, int32_t max_seg, int32_t *out_lens, void *stream);
/* The same from probabilities the caller already holds (records produced with `return_logits`): probs (n, c, t) float32, host or
 * device; lens (n, or NULL = t): valid frames per line, as `record.logits` is cut (rpred.py:200).  No model handle: runs on `device`. */
int  kb_forced_align_probs(const float *probs, int probs_on_device, int32_t n, int32_t c, int32_t t, const int32_t *lens,
                           const int32_t *tokens, const int32_t *tok_off, int32_t *seg_token, int32_t *seg_start, int32_t *seg_end,
                           float *seg_score, int32_t *seg_counts, int32_t max_seg, int device, void *stream);

/* ---- bbox line extraction + the PIL half of the input transforms on the device ----------------------------
 * Replaces, for bbox lines of horizontal text, `im.crop(box)` (kraken/lib/segmentation.py:1631-1643) and the image half of
 * ImageInputTransforms (kraken/lib/dataset/utils.py:123-147): Grayscale, `pil_fixed_resize` = img.resize((int(w * oh / h), oh), LANCZOS)
 * (kraken/lib/functional_im_transforms.py:58-82) and v2.Pad(pad, fill=255) - bit-identical to Pillow's 8-bit resampler (the fixed-point
 * coefficient windows are computed on the host through the same libm calls Pillow makes; the integer convolutions run on the GPU).
 * page:  page_h x page_w x channels uint8, channels = 1 ('L') or 3 ('RGB', interleaved as PIL / numpy hold it), host or device.
 * boxes: n x (x0, y0, x1, y1), inside the page.  out_h: the model's input height.  pad: white columns on both ends.
 * lines: n x 1 x out_h x wmax uint8 on the DEVICE (caller allocated); line i occupies columns [0, widths[i]), the rest of its rows is
 *        left untouched.  widths (host, n): int(w * out_h / h) + 2 * pad.  invert_max (host, n, may be NULL): `im.max()` of
 *        tensor_invert as 0..255 (255 whenever pad > 0).  Hand lines / widths / invert_max to kb_recognize_u8 or
 *        kb_recognize_async(KB_DTYPE_U8, lines_on_device = 1) on the same stream.
 * kb_line_width: the padded width one box will have (to size wmax); 0 = the reference's resize raises for such a box.
 * Not covered (stay in kraken): baseline / polygon line extraction, the centre normaliser of legacy bbox models (valid_norm), vertical text. */
int32_t kb_line_width(int32_t box_w, int32_t box_h, int32_t out_h, int32_t pad);
int  kb_prepare_lines_u8(kb_model *m, const uint8_t *page, int page_on_device, int32_t page_h, int32_t page_w, int32_t channels,
                         int32_t n, const int32_t *boxes, int32_t out_h, int32_t pad, uint8_t *lines, int32_t wmax,
                         int32_t *widths, int16_t *invert_max, void *stream);

/* ---- asynchronous pipeline ------------------------------------------------------------------------
 * The calls above finish with the results in the caller's arrays (one cudaStreamSynchronize each).  A serving loop that wants the
 * copies, the host work and the kernels of consecutive batches to overlap uses ONE model handle with `depth` pipeline slots instead
 * (own stream, activation arena and pinned result block each; ONE copy of the weights): kb_recognize_async enqueues a batch into the
 * next slot and returns a ticket without waiting for the device; kb_wait(ticket) sleeps until that batch is done (blocking-sync
 * event, no spinning) and unpacks its labels.  One host thread keeps `depth` batches in flight:
 *
 *     kb_set_pipeline_depth(m, 4);
 *     for (i = 0; i < nbatches; ++i) {
 *         if (i >= 4) kb_wait(m, t[i - 4], ...);                        // oldest first: its slot is the next to be reused
 *         kb_recognize_async(m, batch[i], KB_DTYPE_F32, 0, n, h, w, widths, NULL, 1.f, T, NULL, &t[i]);
 *     }
 *
 * lines: float32 (KB_DTYPE_F32, kb_recognize semantics) or uint8 (KB_DTYPE_U8, kb_recognize_u8 semantics incl. invert_max).  A HOST
 * `lines` buffer must stay valid and unchanged until kb_wait returns (pin it for a truly asynchronous copy); `widths` / `invert_max`
 * are copied during the call.  For a DEVICE buffer `input_stream` is the stream it was produced on (the slot's stream waits for the
 * work queued there so far).  A batch whose activations left the fp16 operand range is repeated on the fp32 kernels inside kb_wait.
 * kb_recognize_async fails with KB_ERR_SPEC when every slot still holds an un-waited ticket.  Replaces the reference's one-batch-at-
 * a-time loop over `_rec_predict` (kraken/lib/vgsl/rpred.py:126-131,171-176,210-229).                                              */
int  kb_set_pipeline_depth(kb_model *m, int32_t depth);       /* 1..16 slots; allowed only while no ticket is in flight */
int  kb_pipeline_depth(const kb_model *m);
int  kb_recognize_async(kb_model *m, const void *lines, int dtype, int lines_on_device, int32_t n, int32_t h, int32_t w,
                        const int32_t *widths, const int16_t *invert_max, float temperature, int32_t max_out, void *input_stream,
                        int64_t *ticket);
/* counts[i] is the number of labels line i decoded to; it can exceed max_out, in which case only the first max_out are stored
 * (the same holds for kb_recognize / kb_recognize_u8 / kb_ctc_greedy_decode).                                                     */
int  kb_wait(kb_model *m, int64_t ticket, int32_t *labels, int32_t *starts, int32_t *ends, float *confs, int32_t *counts,
             int32_t *out_lens);

/* ---- decoder hook -------------------------------------------------------------------------------
 * replaces kraken.lib.ctc_decoder.greedy_decoder (ctc_decoder.py:35-72) for a (N, C, W) probability
 * tensor; lens: n int32 (NULL => all W, as the reference allows for N == 1).                        */
int  kb_ctc_greedy_decode(const float *probs, int probs_on_device, int32_t n, int32_t c, int32_t w,
                          const int32_t *lens, int32_t *labels, int32_t *starts, int32_t *ends,
                          float *confs, int32_t *counts, int32_t max_out, int device, void *stream);

/* ---- segmentation forward -----------------------------------------------------------------------
 * replaces nn(page) -> F.interpolate(o, size) -> sigmoid (spred.py:268-272, blla.py:121-125), batched.
 * pages: n x C x h x w.  heatmap: n x C' x out_h x out_w fp32.                                      */
int  kb_segment(kb_model *m, const float *pages, int pages_on_device, int32_t n, int32_t h, int32_t w,
                int32_t out_h, int32_t out_w, float *heatmap, int heatmap_on_device, void *stream);

/* ---- introspection for tests / profiling -------------------------------------------------------- */
/* output of leaf layer `name` from the most recent forward on this handle, as NCHW host fp32.
 * dims_only != 0: only fills dims.  Valid until the next call on the handle.                        */
int  kb_debug_layer_output(kb_model *m, const char *name, int32_t dims[4], float *out_host, int dims_only);
/* C[M][N] = A[M][K] * B[N][K]^T + bias through the engine's GEMM kernels (use_tc: 1 = tcgen05 split-fp16 kernel,
 * 0 = CUDA-core fp32 kernel); host pointers.  Unit-test hook for the kernels behind Linear / LSTM projection.   */
int  kb_debug_gemm(const float *a, const float *b, const float *bias, float *c, int32_t M, int32_t N, int32_t K,
                   int use_tc, int device);
/* the host half of kb_prepare_lines_u8: Pillow's fixed-point LANCZOS windows of one axis (Resample.c precompute_coeffs +
 * normalize_coeffs_8bpc).  *ksize = window capacity; bounds [out_size][2] = (first source index, count); kk [out_size][*ksize].
 * bounds / kk may be NULL to query ksize only.  Host only (usable without a GPU): lets the CPU tests pin the tables against Pillow. */
int  kb_debug_axis_coeffs(int32_t in_size, int32_t out_size, int32_t *ksize, int32_t *bounds, int32_t *kk, int32_t kk_cap);
/* number of kernels this handle launched since creation / last reset (bench `gpu_launches`) */
int64_t kb_launch_count(const kb_model *m);
void    kb_reset_launch_count(kb_model *m);
/* The tensor-core layers read activations as two fp16 operand planes (22 significand bits, |x| <= 65504).  A call in
 * which an activation left that range is transparently repeated on the fp32 CUDA-core kernels; this counts those
 * repeats since creation (0 for every sane model; a persistent non-zero rate means KB_GEMM=ffma is the better mode). */
int64_t kb_range_fallback_count(const kb_model *m);
/* device-side stage timing of the most recent compute call on this handle.  With kb_set_timing(m, 1) every
 * stage (one per leaf layer; LSTMs as "<name>.xproj" + "<name>.rec"; "stage_in", "decode", "emit",
 * "upsample_sigmoid") is bracketed by CUDA events on the launching stream.  kb_timing_count() returns the number
 * of stages recorded, kb_timing_entry() name + milliseconds of one of them (in execution order).            */
int  kb_set_timing(kb_model *m, int enabled);
int  kb_timing_count(kb_model *m);
int  kb_timing_entry(kb_model *m, int index, char *name, size_t cap, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* KRAKEN_B200_H */
