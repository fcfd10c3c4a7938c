/*
 * mural_b200.h — C ABI of the B200-native MuRaL hot path (libmural_b200.so).
 *
 * The reference (CaiLiLab/MuRaL v1.2.0) is pure Python and has no FFI; its drop-in seams are three
 * Python call sites.  Every entry point below names the reference function(s) whose
 * work it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stubs a
 * MuRaL maintainer would add at those call sites.
 *
 * Conventions
 *  - plain C types only; `d_*` arguments are CUDA device pointers, `h_*` host pointers;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); nothing here
 *    synchronises the device unless the name ends in `_host` (those copy H2D/D2H and sync `stream`);
 *  - every function returns 0 on success, non-zero on error; mural_last_error() gives the message of
 *    the last failure on the calling thread.  Nothing calls exit() (the reference sys.exit()s);
 *  - a *site* is (pos, meta): pos = 0-based BED start on its chromosome (int32), meta packs
 *    strand | label<<1 | chrom_index<<8   (strand: 0 '+', 1 '-').
 */
#ifndef MURAL_B200_H
#define MURAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MURAL_ABI_VERSION 2

#define MURAL_MODEL_SNV 0
#define MURAL_MODEL_INDEL 1

#define MURAL_META(strand, label, chrom) ((int32_t)(((chrom) << 8) | (((label) & 0x7f) << 1) | ((strand) & 1)))

/* compute modes of the conv stack */
#define MURAL_MODE_FP32 0 /* CUDA-core fp32 FMA; the "fp32-equivalent" mode (1e-3 gate)           */
#define MURAL_MODE_BF16 1 /* tcgen05 implicit GEMM, bf16 operands, fp32 accumulate (5e-3 gate)   */
#define MURAL_MODE_AUTO 2 /* product default: BF16 for every site + the fp32-equivalent path again for the sites whose
                             expanded window holds a non-ACGT symbol or overhangs its chromosome (list built on device) */

const char* mural_last_error(void);
int mural_abi_version(void);
/* number of kernel launches issued by this library since the last mural_reset_launch_count() */
int64_t mural_launch_count(void);
void mural_reset_launch_count(void);
/* optional per-kernel CUDA-event profile: between begin and end every launch of this library is bracketed
 * by two events on its stream; end() synchronises the device and writes
 * {"<kernel>": {"count": n, "ms": total}, ...} as JSON into buf; returns the number of launches seen. */
void mural_profile_begin(void);
int64_t mural_profile_end(char* buf, int64_t cap);

/* ------------------------------------------------------------------------------------------------
 * Packed reference genome (replaces SeqIO.to_dict(...) + the python `str` genome,
 * MuRaL/data/preprocessing.py:836, and the per-character dict lookups at :698-700, :809-813).
 *
 * Layout in HBM: 2 bits/base (A0 C1 G2 T3, 16 bases per uint32, little-endian within the word),
 * 1 bit/base "not ACGT" mask (32 bases per uint32) and a sorted run table
 * (global start, global end, symbol 4..14 = R Y M S W K B D H V N) for the masked bases.
 * Chromosomes are concatenated, each starting on a 64-base boundary.
 * Characters are case-folded; anything outside ACGTRYMSWKBDHVN is an error (the reference raises
 * KeyError on it).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_genome mural_genome_t;

int mural_genome_create(int32_t n_chrom, const char* const* h_seqs, const int64_t* h_lens, int device,
                        mural_genome_t** out);
void mural_genome_destroy(mural_genome_t* g);
int32_t mural_genome_n_chrom(const mural_genome_t* g);
int64_t mural_genome_chrom_len(const mural_genome_t* g, int32_t chrom);
int64_t mural_genome_device_bytes(const mural_genome_t* g);
int64_t mural_genome_n_exception_runs(const mural_genome_t* g);
/* host copy of the non-ACGT runs (R Y M S W K B D H V N) as chromosome index / start / end (end exclusive), sorted; arrays
 * of mural_genome_n_exception_runs() elements.  Lets a caller tell which site windows contain such symbols. */
int mural_genome_exception_runs(const mural_genome_t* g, int32_t* h_chrom, int64_t* h_start, int64_t* h_end);

/* ------------------------------------------------------------------------------------------------
 * Encoders (bit-exact with the reference; exposed for parity tests and for callers that still want
 * the reference tensors).
 *
 * mural_encode_local  : seq_digit_encoder + process_local_seq_* (preprocessing.py:636-723, 479-522)
 *                       -> int64 [n, 2R+1-(k-1)] (snv) / [n, 2R-(k-1)] (indel)
 * mural_encode_onehot : seq_ohe_encoder + distal_encoding_by_region (preprocessing.py:756-816, 978-999)
 *                       -> float32 [n, 4, W], W = 2R+1 (snv) / 2R (indel)
 * ---------------------------------------------------------------------------------------------- */
int mural_encode_local(const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                       int32_t radius, int32_t order, int32_t model_type, int64_t* d_out, void* stream);
int mural_encode_onehot(const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                        int32_t radius, int32_t model_type, float* d_out, void* stream);
int mural_encode_local_host(const mural_genome_t* g, const int32_t* h_pos, const int32_t* h_meta, int64_t n,
                            int32_t radius, int32_t order, int32_t model_type, int64_t* h_out);
int mural_encode_onehot_host(const mural_genome_t* g, const int32_t* h_pos, const int32_t* h_meta, int64_t n,
                             int32_t radius, int32_t model_type, float* h_out);
/* one-hot tensor -> symbol codes (0..14, 255 = not a reference one-hot column); lets the drop-in
 * Network2.forward(distal_x) accept the reference's own input tensors. */
int mural_onehot_to_symbols(const float* d_onehot, int64_t n, int32_t W, uint8_t* d_sym, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MuRaL-snv network (Network2, MuRaL/model/model_snv.py:290-525).
 *
 * Parameters are exchanged as ONE flat fp32 blob whose layout is queried by name: entry i has the
 * reference state_dict key (e.g. "RBs1_2.0.conv1.weight"), an element offset and a size.  The
 * aliased "*.layer.N.*" duplicates of the reference state_dict are not part of the blob.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_snv_config {
  int32_t local_radius;   /* config['local_radius']                      */
  int32_t local_order;    /* config['local_order'] (k of the k-mer)      */
  int32_t distal_radius;  /* config['distal_radius'] -> L = 2R+1         */
  int32_t hidden1;        /* config['local_hidden1_size']                */
  int32_t hidden2;        /* config['local_hidden2_size']                */
  int32_t channels;       /* config['CNN_out_channels']                  */
  int32_t kernel_size;    /* config['CNN_kernel_size'] (conv1/2/3)       */
  int32_t n_class;        /* config['n_class']                           */
  int32_t n_cont;         /* common_model_config['n_cont']: continuous (bigWig mean) features of the local branch
                             (first_bn_layer + wider first Linear, model_snv.py:326-334,457-463); 0 for every shipped model */
} mural_snv_config_t;

typedef struct mural_snv_model mural_snv_model_t;

int mural_snv_model_create(const mural_snv_config_t* cfg, int device, mural_snv_model_t** out);
void mural_snv_model_destroy(mural_snv_model_t* m);
int32_t mural_snv_model_n_tensors(const mural_snv_model_t* m);
/* is_buffer: 0 = trainable parameter, 1 = BatchNorm running statistic.  All trainable tensors come
 * first in the blob, so blob[0 : n_trainable) is the flat parameter vector optimisers work on. */
int mural_snv_model_tensor(const mural_snv_model_t* m, int32_t i, const char** name, int64_t* offset,
                           int64_t* numel, int32_t* is_buffer);
int64_t mural_snv_model_n_params(const mural_snv_model_t* m);    /* blob length in floats */
int64_t mural_snv_model_n_trainable(const mural_snv_model_t* m); /* leading trainable part   */
/* Load eval-mode weights (BatchNorm running stats included in the blob); folds and uploads. */
int mural_snv_model_load(mural_snv_model_t* m, const float* h_blob, int64_t n);

/* n_cont > 0 only: the continuous features cont_x [n, n_cont] (float32, device; row i <-> site i) of the NEXT forward call
 * (local_input[0] of Network2.forward, model_snv.py:448,457-463).  The pointer is consumed by that call — mural_snv_forward* or
 * mural_snv_train_forward (which keeps reading it until the matching mural_snv_train_backward has run: first_bn_layer's
 * gradients).  Models with continuous features run in MURAL_MODE_FP32. */
int mural_snv_set_cont(mural_snv_model_t* m, const float* d_cont);

/* model_predict_m body (MuRaL/model/nn_utils.py:48-65) for n sites: gather + Network2.forward (eval).
 * d_logp: float32 [n, n_class] log-probabilities, exactly what Network2.forward returns. */
int mural_snv_forward(mural_snv_model_t* m, const mural_genome_t* g, const int32_t* d_pos,
                      const int32_t* d_meta, int64_t n, int32_t mode, float* d_logp, void* stream);
/* Same network on the reference's own tensors (cat_x int64 [n,n_cat], distal_x float32 [n,4,L]). */
int mural_snv_forward_tensors(mural_snv_model_t* m, const int64_t* d_cat, const float* d_distal, int64_t n,
                              int32_t L, int32_t mode, float* d_logp, void* stream);
/* End-to-end: host site arrays in, host log-probs out (H2D + kernels + D2H + stream sync). */
int mural_snv_predict_host(mural_snv_model_t* m, const mural_genome_t* g, const int32_t* h_pos,
                           const int32_t* h_meta, int64_t n, int32_t mode, float* h_logp, void* stream);
/* CrossEntropyLoss(reduction='sum') over d_logp with labels from meta (nn_utils.py:64): adds into *d_loss */
int mural_ce_sum(const float* d_logp, const int32_t* d_meta, int64_t n, int32_t n_class, double* d_loss,
                 void* stream);
/* 1 when the tcgen05 (MURAL_MODE_BF16) path is compiled in and supports this model's shape */
int mural_snv_tc_available(const mural_snv_model_t* m);
/* number of sites the last MURAL_MODE_AUTO forward recomputed in the fp32-equivalent path */
int64_t mural_snv_last_auto_sites(const mural_snv_model_t* m);
/* sites per workspace chunk of the forward (0 = default); parity-test switch for debug taps */
int mural_snv_set_chunk(mural_snv_model_t* m, int64_t chunk_sites);
int mural_snv_set_debug(mural_snv_model_t* m, int32_t flags); /* bit0: keep taps, bit1: force the generic stem kernel */
/* debug/parity taps: copies an intermediate activation of the LAST forward chunk to the host.
 * name in {"pool1","pool1_2","rb1_2","conv2_2","rb2_2","gmax","gmax_2","logit_local","logit_mid","logit_large"} */
int mural_snv_debug_tap(mural_snv_model_t* m, const char* name, float* h_out, int64_t max_floats,
                        int64_t* n_written);

/* Parity hooks for the tensor-core conv kernels of the fp32-equivalent / training paths (snv_conv_mma.cu), C = 32, ks = 3:
 * one BatchNorm(eval affine a, b) -> Conv1d(32,32,3,padding=1) layer of Network2 (model_snv.py:794-812) on device tensors
 * in the [n*L, 32] row layout, and its weight gradient.  impl: 0 = fp32 FMA kernel, 1 = split-bf16 MMA (two levels),
 * 2 = split-bf16 MMA (three levels, training).  Wt is [tap][ci][co]; dW is [co][ci][tap] (+=), dbias [co] (+=). */
int mural_conv32_layer(const float* d_in, float* d_out, const float* d_res1, const float* d_res2, int64_t n, int32_t L,
                       const float* d_Wt, const float* d_bias, const float* d_a, const float* d_b, int32_t relu_in,
                       int32_t relu_out, int32_t impl, void* stream);
int mural_conv32_wgrad(const float* d_x, const float* d_dy, int64_t n, int32_t L, int32_t relu_in, const float* d_a,
                       const float* d_b, float* d_dW, float* d_dbias, int32_t impl, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MuRaL-indel network (UNet_Small, MuRaL/model/model_indel.py:21-176), eval forward.
 * Window of a site: [start - R + 1, start + 1 + R), length 2R (extend_interval, preprocessing.py:559-567).
 * d_out: float32 [n, n_class] Softplus activations, exactly what UNet_Small.forward returns.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_indel_config {
  int32_t distal_radius; /* config['distal_radius'] -> L = 2R                     */
  int32_t channels;      /* config['CNN_out_channels'] (level i has (i+1)*channels) */
  int32_t kernel_size;   /* config['CNN_kernel_size']                              */
  int32_t n_class;       /* config['n_class']                                      */
  int32_t downsize[6];   /* config['down_list']                                    */
  int32_t use_reverse;   /* config.get('use_reverse', False)                       */
} mural_indel_config_t;
typedef struct mural_indel_model mural_indel_model_t;
int mural_indel_model_create(const mural_indel_config_t* cfg, int device, mural_indel_model_t** out);
void mural_indel_model_destroy(mural_indel_model_t* m);
int32_t mural_indel_model_n_tensors(const mural_indel_model_t* m);
int mural_indel_model_tensor(const mural_indel_model_t* m, int32_t i, const char** name, int64_t* offset, int64_t* numel,
                             int32_t* is_buffer);
int64_t mural_indel_model_n_params(const mural_indel_model_t* m);
int mural_indel_model_load(mural_indel_model_t* m, const float* h_blob, int64_t n);
int mural_indel_forward(mural_indel_model_t* m, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                        float* d_out, void* stream);
/* same network on the reference's own input tensor distal_x float32 [n, 4, 2R] */
int mural_indel_forward_tensors(mural_indel_model_t* m, const float* d_distal, int64_t n, int32_t L, float* d_out, void* stream);
/* Kernel family of the eval forward.  mode 0 (default): one fused tensor-core kernel per U-Net level (lconv + ConvBlock
 * [+ skip, + out_conv + position max], model_indel.py:6-19,158-173; split-bf16 products, fp32 accumulation — fp32-equivalent,
 * gate 1e-3 of the output scale) whenever the shapes allow it (CNN_out_channels a multiple of 8, widest level <= 48 channels);
 * mode 1: the fp32 CUDA-core kernels (one per convolution).  mural_indel_tc_available: 1 if mode 0 runs the level kernels. */
int mural_indel_set_mode(mural_indel_model_t* m, int32_t mode);
int mural_indel_tc_available(const mural_indel_model_t* m);

/* ------------------------------------------------------------------------------------------------
 * Training step (MuRaL/training.py:404-452): train-mode forward (batch-statistic BatchNorm with running-stat
 * update, dropout), backward, and the fused gradient-clip + optimizer update.
 *
 * d_blob   : the model's flat fp32 buffer on the device, layout of mural_snv_model_tensor() (trainable part first,
 *            BatchNorm running statistics after; the forward updates the running statistics in place).
 * d_grads  : flat fp32 gradient buffer, n_trainable floats, same offsets; overwritten by backward.  In data-parallel
 *            training this is the single buffer to all-reduce (sum) between backward and mural_optimizer_step.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_snv_train mural_snv_train_t;
int mural_snv_train_create(mural_snv_model_t* m, mural_snv_train_t** out);
void mural_snv_train_destroy(mural_snv_train_t* t);
/* dropout probabilities (config emb_dropout, local_dropout, distal_fc_dropout; model_snv.py:338-339,385,427) */
int mural_snv_train_set_dropout(mural_snv_train_t* t, float p_emb, float p_local, float p_fc, uint64_t seed);
/* forward of one batch (n >= 2 sites); d_logp float32 [n, n_class] = what Network2.forward returns in train() mode */
int mural_snv_train_forward(mural_snv_train_t* t, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                            int64_t n, float* d_blob, float* d_logp, void* stream);
/* backward of the last forward given dLoss/dlogp [n, n_class] */
int mural_snv_train_backward(mural_snv_train_t* t, const float* d_blob, const float* d_dlogp, float* d_grads, void* stream);
/* CrossEntropyLoss(reduction='sum') on log-probs (training.py:327,425): adds the loss into *d_loss (may be NULL) and
 * writes dLoss/dlogp */
int mural_ce_sum_grad(const float* d_logp, const int32_t* d_meta, int64_t n, int32_t n_class, double* d_loss, float* d_dlogp,
                      void* stream);
/* clip_grad_norm_(max_norm) + optimizer.step() (training.py:434-436, 346-357) over a flat buffer, no host sync.
 * kind 0 Adam (coupled L2 weight decay), 1 AdamW(amsgrad=True), 2 SGD(momentum 0.98, nesterov).  Gradients are
 * multiplied by grad_scale first (1/world_size to average a summed all-reduce, or 1).  step counts from 1.
 * d_scratch: >= 8 bytes of device scratch (receives the squared gradient norm). */
int mural_optimizer_step(int32_t kind, float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_vmax, int64_t n,
                         float lr, float weight_decay, int64_t step, float max_norm, float grad_scale, double* d_scratch,
                         void* stream);
/* Same step with the per-step hyper-parameters on the device, so that the whole training step can be captured once in a
 * CUDA graph and replayed: *d_lr is read by the kernel, *d_step (int64, starts at 0) is incremented by the call before use. */
int mural_optimizer_step_dev(int32_t kind, float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_vmax, int64_t n,
                             const float* d_lr, float weight_decay, int64_t* d_step, float max_norm, float grad_scale,
                             double* d_scratch, void* stream);

/* MuRaL-indel training step (UNet_Small in train() mode, model_indel.py:6-176; loop body training.py:404-452).
 * Same protocol as the snv step: d_blob is the flat fp32 parameter buffer in the layout of mural_indel_model_tensor()
 * (trainable tensors first, BatchNorm running statistics after them; running statistics are updated in place by forward),
 * d_out float32 [n, n_class] = the Softplus activations UNet_Small.forward returns, the loss is CrossEntropyLoss(sum) on them
 * (mural_ce_sum_grad), d_grads is a flat buffer of mural_indel_model_n_params() floats (same offsets; the running-statistic
 * part stays zero) whose first n_trainable floats go to the all-reduce and mural_optimizer_step(_dev).  p_fc: Dropout of
 * out_fc (0.1 in the reference, model_indel.py:147). */
int mural_indel_model_config(const mural_indel_model_t* m, mural_indel_config_t* out);
typedef struct mural_indel_train mural_indel_train_t;
int mural_indel_train_create(mural_indel_model_t* m, mural_indel_train_t** out);
void mural_indel_train_destroy(mural_indel_train_t* t);
int mural_indel_train_set_dropout(mural_indel_train_t* t, float p_fc, uint64_t seed);
int mural_indel_train_forward(mural_indel_train_t* t, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                              int64_t n, float* d_blob, float* d_out, void* stream);
/* same on the reference's own input tensor distal_x float32 [n, 4, 2R] */
int mural_indel_train_forward_tensors(mural_indel_train_t* t, const float* d_distal, int64_t n, int32_t L, float* d_blob,
                                      float* d_out, void* stream);
int mural_indel_train_backward(mural_indel_train_t* t, float* d_blob, const float* d_dout, float* d_grads, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Calibration epilogue (run_predict.py:214-225): softmax(logp) -> FullDirichlet apply
 * (dirichletcal/calib/fulldirichlet.py:78-80, multinomial.py:60-64,235-244) -> optional Poisson
 * calibration (MuRaL/model/calibration.py:10-23).  h_weights: fp64 [k, k+1] or NULL.  Output fp64 [n,k].
 * ---------------------------------------------------------------------------------------------- */
int mural_calibrate(const float* d_logp, int64_t n, int32_t n_class, const double* h_weights,
                    int32_t poisson, double* d_prob, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation metrics of the reference's Evaluator (MuRaL/evaluation/evaluation.py, called every epoch from
 * MuRaL/training.py:488-520) as device-side grouped reductions.  Labels are read from d_meta (MURAL_META), probabilities
 * are fp64 [n, n_class] (mural_calibrate output).  All sums are 64-bit integers: counts exactly, probabilities in fixed
 * point with MURAL_METRIC_SCALE units per 1.0, so tables are reproducible bit for bit; n <= 2^27 per call.
 * mural_kmer_group_stats  the groupby(...).mean() of freq_kmer_comp_multi (:48-67) and, with region_size > 0, of every
 *                   region of evaluate_regional_score (:545-566; calc_avg_prob :196-203 is the sum over a region's groups).
 *                   d_flank: int64 [n, n_cols] order-1 local codes 0..4 (columns us_R..us1, mid, ds1..ds_R of data_local);
 *                   group index = codes of us_d..us1, ds1..ds_d (d = k/2) as a base-5 number, first column most significant
 *                   (= pandas group order).  region_size 0: one region of n sites; otherwise n / region_size regions of
 *                   consecutive sites (the remainder is ignored, as the reference's iloc slices do).
 *                   d_table: int64 [n_regions, 5^(2d), 1 + 2*n_class] = sites, sites per label, fixed-point prob sums.
 * mural_window_runs corr_calc_sub (:124-193): runs of consecutive sites (in d_order if given, else as stored) that share
 *                   (chrom, start // window).  *h_n_runs receives the number of runs (host sync).  With d_rows == NULL only
 *                   counts; otherwise fills int64 [n_runs, 1 + 2*n_class] (same columns as above; n_runs <= max_runs).
 * ---------------------------------------------------------------------------------------------- */
#define MURAL_METRIC_SCALE 68719476736.0 /* 2^36 */
int mural_kmer_group_stats(const int64_t* d_flank, int64_t n, int32_t n_cols, int32_t k, const int32_t* d_meta,
                           const double* d_prob, int32_t n_class, int64_t region_size, int64_t* d_table, void* stream);
int mural_window_runs(const int32_t* d_meta, const int32_t* d_start, const int64_t* d_order, const double* d_prob, int64_t n,
                      int32_t n_class, int32_t window, int64_t* h_n_runs, int64_t* d_rows, int64_t max_runs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host ingest (no GPU involved): streaming BED / FASTA readers, plain or gzip.
 * mural_bed_read   replaces iterating `BedTool(file)` for .chrom/.start/.stop/.score/.strand
 *                  (MuRaL/data/preprocessing.py:39-106, 752-754): BED6, score = label, strand '+' -> 0 else 1;
 *                  chromosome indices in order of first appearance.
 * mural_fasta_read replaces `SeqIO.to_dict(SeqIO.parse(open(ref_genome), 'fasta'))` (preprocessing.py:836): records in
 *                  file order, id = first word of the header; a duplicate id is an error ("ValueError: Duplicate key").
 *                  mural_fasta_seq pointers stay valid until mural_fasta_destroy and can be passed to
 *                  mural_genome_create directly.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_bed mural_bed_t;
typedef struct mural_fasta mural_fasta_t;
int mural_bed_read(const char* path, mural_bed_t** out);
int64_t mural_bed_n(const mural_bed_t* b);
int32_t mural_bed_n_chrom(const mural_bed_t* b);
const char* mural_bed_chrom_name(const mural_bed_t* b, int32_t i);
int mural_bed_columns(const mural_bed_t* b, int32_t* chrom, int64_t* start, int64_t* end, int8_t* strand, int64_t* label);
void mural_bed_destroy(mural_bed_t* b);
/* Emission order of bed_reader (MuRaL/data/preprocessing.py:39-106): windows of `segment_center` bp anchored at the first site of
 * the first chromosome block (at 1 for later blocks), '+' batch before '-' batch inside each window, file order inside a batch.
 * perm [n]: file index of the k-th emitted site; batch_sizes [capacity n]: sizes of the non-empty batches, *n_batches of them. */
int mural_segment_order(const int32_t* chrom, const int64_t* start, const int8_t* strand, int64_t n, int64_t segment_center,
                        int64_t* perm, int64_t* batch_sizes, int64_t* n_batches);
/* Site records in emission order from the BED columns in file order (what CombinedDatasetNP keeps per sample,
 * MuRaL/data/preprocessing.py:850-954, reduced to 8 bytes per site): pos = start[perm], strand, label,
 * chrom = chrom_map[chrom[perm]] (genome index of the BED's i-th chromosome) and meta = MURAL_META(strand, label, chrom).
 * A label outside [0, 127] is an error. */
int mural_pack_sites(const int64_t* perm, int64_t n, const int32_t* chrom, const int64_t* start, const int8_t* strand,
                     const int64_t* label, const int64_t* chrom_map, int32_t n_chrom, int32_t* pos_out, int8_t* strand_out,
                     int64_t* label_out, int64_t* chrom_out, int32_t* meta_out);
int mural_fasta_read(const char* path, mural_fasta_t** out);
int32_t mural_fasta_n(const mural_fasta_t* f);
const char* mural_fasta_name(const mural_fasta_t* f, int32_t i);
const char* mural_fasta_seq(const mural_fasta_t* f, int32_t i);
int64_t mural_fasta_len(const mural_fasta_t* f, int32_t i);
void mural_fasta_destroy(mural_fasta_t* f);

/* ------------------------------------------------------------------------------------------------
 * bigWig tracks (replaces pyBigWig in get_mean_bw_for_bed, MuRaL/data/preprocessing.py:725-750): host reader.
 * window_means: out[i] = mean(nan_to_num(values(chrom, max(lo[i], 0), min(hi[i], chrom length)))), end exclusive; bases
 * without data count as 0, an empty window gives NaN (np.mean of an empty array).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mural_bigwig mural_bigwig_t;
int mural_bigwig_open(const char* path, mural_bigwig_t** out);
int32_t mural_bigwig_n_chrom(const mural_bigwig_t* b);
const char* mural_bigwig_chrom_name(const mural_bigwig_t* b, int32_t i);
int64_t mural_bigwig_chrom_len(const mural_bigwig_t* b, int32_t i);
int mural_bigwig_window_means(mural_bigwig_t* b, const char* chrom, int64_t n, const int64_t* lo, const int64_t* hi, double* out);
void mural_bigwig_close(mural_bigwig_t* b);

/* ------------------------------------------------------------------------------------------------
 * Prediction TSV (run_predict.py:228-239): pred_df.to_csv(pred_file, sep='\t', float_format='%.4g', index=False) with
 * columns chrom, start, end, strand, mut_type, prob0..prob{k-1}.  Host arrays, rows already in output order
 * (sorted by chrom name, start); strand is one char per row; formatted by n_threads host threads.
 * ---------------------------------------------------------------------------------------------- */
int mural_write_tsv(const char* path, int64_t n, int32_t n_class, const char* const* chrom_names, const int32_t* chrom_idx,
                    const int64_t* start, const int64_t* end, const char* strand, const double* mut_type, const double* prob,
                    int32_t n_threads);
/* the writer's "%.4g" formatter on one value (parity hook: tests compare it with printf); out32: >= 32 bytes, NUL-terminated */
int mural_format_g4(double v, char* out32);

#ifdef __cplusplus
}
#endif
#endif /* MURAL_B200_H */
