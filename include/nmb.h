/*
 * nmb.h -- C ABI of libnmb.so: the B200 (sm_100a) implementation of the ONE hot path of
 * soz223/multi_modal_normative_modeling: training many small conditional VAEs at once and
 * scoring per-subject / per-ROI deviations.
 *
 * The reference has no FFI layer: its boundary is a Python class + CLI contract
 * (SURVEY.md section 8b).  Each entry point below names the reference interface
 * (file:line, relative to the upstream repository) whose work it replaces; the Python
 * nn.Module shims in multi_modal_normative_modeling_b200/ bind these through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - every function returns 0 on success, non-zero on error; nmb_last_error() returns
 *     a thread-local message for the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller (torch tensors) unless a
 *     parameter is documented as host memory.  All tensors are fp32, row-major.
 *   - every launch takes the CUDA stream as a void* (cudaStream_t); no internal threads,
 *     no hidden synchronisation except where documented.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NMB_H_
#define NMB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMB_MAX_MOD 16    /* modalities per model (HCP uses 12 + early fusion) */
#define NMB_MAX_HIDDEN 4  /* hidden layers per encoder/decoder (hz_para_list up to "1024 512 256 32") */
#define NMB_MAX_HEAD 3    /* hidden layers of a supervised head (the regressor has 2: 128, 64) */

/* combine_latent kinds, cVAE.py:1144-1164 (case-insensitive strings in the reference) */
enum { NMB_COMBINE_POE = 0, NMB_COMBINE_GPOE = 1, NMB_COMBINE_MOE = 2, NMB_COMBINE_MOPOE = 3 };
/* reconstruction term: Gaussian log-likelihood with learned logvar_out (cVAE.py:14-15,193-206)
 * or -MSE(mean) (multimodal_kfold_cvae_nmmlp.py:124-127) */
enum { NMB_LOSS_GAUSS_LL = 0, NMB_LOSS_NEG_MSE = 1 };
/* supervised head on top of the multimodal cVAE.  NMB_HEAD_REGRESSION = cVAE_multimodal_regression
 * (cVAE.py:2211-2347): regressor MLP Linear(sum D, h0) ReLU ... Linear(h_last, 1) on the concatenated residuals
 * x_m - x_recon_m.loc (:2321-2325), loss total += head_weight * MSE(fi_pred, true_fi) (:2334-2347).
 * NMB_HEAD_ENDTOEND = cVAE_multimodal_endtoend v2 (cVAE.py:2004-2207, driven by multimodal_kfold_cvae_nmpmcont.py): shared
 * encoders, PoE fusion (:2081-2088), TWO decoder sets (health / disease), a classifier Linear -> BatchNorm1d -> ReLU ->
 * Dropout ... -> Linear(., 2) on z (:2004-2018), loss = w_rec (rec_health + rec_disease) + w_kl kl + cross-entropy +
 * w_con * contrastive hinge on the per-subject deviations of the two decoder sets (:2131-2196).  n_mod <= 8. */
enum { NMB_HEAD_NONE = 0, NMB_HEAD_REGRESSION = 1, NMB_HEAD_ENDTOEND = 2 };
/* indices into NmbArch.head_params for NMB_HEAD_ENDTOEND (loss_function defaults :2131: margin 1, 0.1, 0.1, 0.1; dropout .5) */
enum { NMB_HP_MARGIN = 0, NMB_HP_W_CONTRASTIVE = 1, NMB_HP_W_KL = 2, NMB_HP_W_REC = 3, NMB_HP_DROPOUT = 4 };

/* Model family.  NMB_FAMILY_DMVAE = the baselines built on VariationalEncoder / VariationalDecoder
 * (cVAE.py:1454-1480): DMVAE (:1491-1618), mmVAEPlus (:1895-2002; the same code with beta = 0.05) and WeightedDMVAE
 * (:1620-1752; learnable per-modality loss weights).  Encoder m: x -> fc1 ReLU -> fc2 ReLU -> (fc_mu, fc_logvar), NO
 * covariates; the first s_dim latent dimensions of every modality are PRIVATE (passed on as their mean), the rest are
 * shared (product of experts over the modalities, reparameterised); decoder m: [z_shared | mu_private_m] -> fc1 ReLU ->
 * fc2 ReLU -> sigmoid(fc_out); loss = sum_m w_m (beta * kl_shared + 0.5 * sum_d (x - x_recon)^2 / B).  With the
 * reference's defaults (s_dim = c_dim = 29 >= latent 10) there is no shared part and the model is M deterministic
 * autoencoders.  Requires n_hidden == 2, c_dim == 0 (packed rows are [x | 1]), head_kind == 0.
 * NMB_FAMILY_MVTCAE = mvtCAE (cVAE.py:1754-1893): the cVAE_multimodal architecture (covariates, logvar_out, alphas) with
 * total = sum_m (kl + 1e-5 * ll_m + beta * tc) -- the log-likelihood enters with a PLUS sign as written (:1862) -- where
 * tc = - sum_i mean_m logsumexp_b mu_m[b, i] (:1846-1853; its first term cancels itself), beta = 1e-4; the fused variance is
 * clamped to >= 1e-6 (:1815) and the 'poe' branch hands VARIANCES to ProductOfExperts2, which exponentiates them again
 * (:1778, :1800): reproduced as written. */
enum { NMB_FAMILY_CVAE = 0, NMB_FAMILY_DMVAE = 1, NMB_FAMILY_MVTCAE = 2 };

/* Architecture of one ensemble member = the constructor arguments of
 * cVAE_multimodal(input_dim_list, hidden_dim, latent_dim, c_dim, ..., modalities, non_linear)
 * (cVAE.py:1088-1116); the single-modality cVAE (cVAE.py:391-411) is n_mod == 1. */
typedef struct {
  int32_t n_mod;
  int32_t input_dims[NMB_MAX_MOD];
  int32_t n_hidden;
  int32_t hidden[NMB_MAX_HIDDEN];
  int32_t latent;
  int32_t c_dim;
  int32_t combine;    /* NMB_COMBINE_* */
  int32_t loss_kind;  /* NMB_LOSS_* */
  int32_t non_linear; /* leaky_relu(0.01) after hidden layers (cVAE.py:166-167) */
  int32_t head_kind;  /* NMB_HEAD_*; members with a head run on the generic engines */
  int32_t n_head_hidden;
  int32_t head_hidden[NMB_MAX_HEAD]; /* 128, 64 (cVAE.py:2248-2255) */
  float head_weight;  /* lambda_reg (cVAE.py:2334; the trainer passes 1.0) */
  float head_params[6]; /* NMB_HP_* (NMB_HEAD_ENDTOEND) */
  int32_t family;     /* NMB_FAMILY_* */
  int32_t s_dim;      /* NMB_FAMILY_DMVAE: private latent dimensions per modality (the reference passes c_dim) */
  int32_t weighted;   /* NMB_FAMILY_DMVAE: 1 = WeightedDMVAE (weights in the NMB_SLOT_ALPHA slots, total = kl - ll) */
  float beta;         /* NMB_FAMILY_DMVAE: total = beta * kl - ll (1.0 DMVAE, 0.05 mmVAEPlus); NMB_FAMILY_MVTCAE: weight of tc (1e-4) */
} NmbArch;

/* One tensor of the reference's state_dict inside the packed per-model parameter buffer.
 * A linear layer (in -> out) is stored as an augmented matrix [out][ld], ld = roundup4(in+1):
 * row n = { weight[n][0..in-1], bias[n], 0... }.  The mean and logvar heads are stacked
 * (rows 0..Z-1 = enc_mean_layer, Z..2Z-1 = enc_logvar_layer). */
enum { NMB_SLOT_ENC = 0, NMB_SLOT_ENC_MEAN = 1, NMB_SLOT_ENC_LOGVAR = 2, NMB_SLOT_DEC = 3,
       NMB_SLOT_DEC_MEAN = 4, NMB_SLOT_LOGVAR_OUT = 5, NMB_SLOT_ALPHA = 6,
       NMB_SLOT_HEAD = 7 /* layer l of the head = regressor.{2l} (cVAE.py:2248-2255) or classifier.classifier.{4l}
                            (cVAE.py:2010-2015); modality 0 */,
       NMB_SLOT_HEAD_BN = 8 /* BatchNorm1d after head layer l = classifier.classifier.{4l+1}: rows = 5 vectors of `cols`
                               entries at stride ld: weight, bias, running_mean, running_var, [num_batches_tracked] */,
       NMB_SLOT_DEC2 = 9, NMB_SLOT_DEC2_MEAN = 10, NMB_SLOT_LOGVAR_OUT2 = 11 /* the second (disease) decoder set */ };
typedef struct {
  int32_t kind;      /* NMB_SLOT_* */
  int32_t modality;
  int32_t layer;     /* index inside encoder_layers / decoder_layers */
  int32_t rows;      /* out features (1 for logvar_out / alpha) */
  int32_t cols;      /* in features  (D for logvar_out, 1 for alpha) */
  int32_t ld;        /* row stride in floats */
  int64_t offset;    /* float offset of element [0][0]; bias of row n is at offset + n*ld + cols */
} NmbSlot;

/* One ensemble member: architecture, training rows, optimiser settings, state buffers.
 * Replaces one iteration of the fold / modality / seed loops of
 * multimodal_kfold_train_cvae_supervised.py:68-212. */
typedef struct {
  NmbArch arch;
  const float* xc[NMB_MAX_MOD]; /* per modality: packed rows [n_rows][ldx] from nmb_pack_rows */
  int32_t n_rows;               /* training rows (after bootstrap + merge) */
  int32_t batch;                /* 256 in the reference (train script :116); no shuffle, last batch partial */
  uint64_t seed;                /* Philox key of the in-kernel eps stream */
  float lr, beta1, beta2, adam_eps; /* torch.optim.Adam defaults 1e-4, .9, .999, 1e-8 (cVAE.py:1111-1116) */
  const float* lr_steps;        /* optional [n_lr_steps] per-step LR (nmmlp cyclic schedule :363-381) */
  float* params;                /* [n_params] packed, see NmbSlot */
  float* adam_m;                /* [n_params] exp_avg    */
  float* adam_v;                /* [n_params] exp_avg_sq */
  float* grads;                 /* [n_params] or NULL; written when NMB_TRAIN_WRITE_GRADS */
  int64_t n_lr_steps;           /* entries of lr_steps; a train call that would step past it FAILS (no silent read
                                   beyond the schedule).  Ignored when lr_steps is NULL. */
  const float* y;               /* head target per training row [n_rows] (FI, ..._regression.py:86-87); NULL without a head */
  const int32_t* row_order;     /* optional [n_order_epochs][n_mod][n_rows]: the dataset row each modality's loader yields
                                   at position i of that epoch -- DataLoader(shuffle=True) draws one permutation PER
                                   MODALITY per epoch (..._regression.py:94, 122); the target follows modality 0 (:125).
                                   NULL = rows in order (shuffle=False).  Generic engines only. */
  int64_t n_order_epochs;       /* epochs row_order covers; training past it FAILS */
  const float* drop_keep;       /* optional injected dropout keep flags (0 / 1) of the classifier, [n_drop_steps][batch][sum of
                                   the head's hidden widths], step s reads entry s mod n_drop_steps (parity tests and the
                                   per-step module API, where torch draws the masks); NULL = in-kernel Philox, stream 2 */
  int64_t n_drop_steps;
} NmbMember;

typedef struct NmbEnsemble NmbEnsemble;

/* ---- errors / introspection ------------------------------------------------------- */
const char* nmb_last_error(void);
int nmb_version(void);
int nmb_device_count(int* count);

/* ---- layout ------------------------------------------------------------------------ */
int nmb_arch_param_count(const NmbArch* arch, int64_t* n_params);
/* Fills up to max_slots entries in reference state_dict order
 * (alpha_m_list.*, encoder_list.*, decoder_list.*; cVAE.py:1107-1109); *n_slots = total. */
int nmb_arch_slots(const NmbArch* arch, NmbSlot* slots, int32_t max_slots, int32_t* n_slots);
/* Row stride of a packed dataset row [x | c | 1 | 0-pad]. */
int nmb_packed_row_stride(int32_t d, int32_t c_dim, int32_t* ldx);

/* ---- data staging ------------------------------------------------------------------ */
/* out[i] = { x[i][0..d), c[i][0..c_dim), 1, 0... }  -- the cat((x, c), dim=1) of
 * Encoder.forward (cVAE.py:163) done once per dataset instead of once per minibatch,
 * plus the constant-1 column that carries the bias through the GEMMs. */
int nmb_pack_rows(const float* x, const float* c, int64_t n_rows, int32_t d, int32_t c_dim,
                  float* out, void* stream);

/* ---- preprocessing prologue of a fold, resident on the GPU ------------------------------------------
 * x: the raw float64 feature table of one modality [n_all][ld] in HBM; idx: row positions of the fold's train or test
 * rows (host-computed: bootstrap + merge order are integer work on the legacy numpy RNG, utils.py:73-168), or NULL for
 * rows 0..n-1.  float64 arithmetic, results bit-identical to the host calls they replace. */

/* RobustScaler().fit (train script :101-102): center[j] = np.nanmedian, scale[j] = q75 - q25 (np.nanpercentile, linear),
 * zeros -> 1.  n <= 8192 rows (shared-memory sort); center / scale: device float64 [d]. */
int nmb_robust_fit(const double* x, int64_t ld, int32_t d, const int32_t* idx, int32_t n, double* center, double* scale,
                   void* stream);
/* v.rank(method='first') -> pd.qcut(q) bin per selected row (train script :105-114).  edges: device float64 [q + 1], the
 * bin edges pandas computes for n ranks (they depend on n only; the host passes pandas' own values).  bins: int32 [n]. */
int nmb_rank_bins(const double* v, const int32_t* idx, int32_t n, const double* edges, int32_t q, int32_t* bins, void* stream);
/* One pass: gather rows, (x - center) / scale in float64 -> fp32, np.eye(n_age)[age_bin] | np.eye(n_sex)[sex_bin], the
 * constant-1 column, zero padding: the packed rows of nmb_pack_rows ([n][roundup4(d + n_age + n_sex + 1)]). */
int nmb_pack_rows_scaled(const double* x, int64_t ld, int32_t d, const int32_t* idx, int64_t n, const double* center,
                         const double* scale, const int32_t* age_bin, int32_t n_age, const int32_t* sex_bin, int32_t n_sex,
                         float* out, void* stream);

/* ---- on-disk contract writer (f2) ---------------------------------------------------- */
/* One table of the test program's CSV families (multimodal_kfold_test_cvae_supervised.py:116-178, DataFrame.to_csv(index=
 * False)): `header` line, then per row `row_prefix[r]` (the leading cells already joined with ',', or NULL) followed by
 * n_cols numbers of the HOST block `body` (float32 or float64, row stride ld), printed exactly as pandas prints a float
 * column (shortest round-trip digits, numpy's positional / scientific rule, nan -> empty).  Byte-identical files; rows are
 * formatted by `threads` host threads (0 = all cores). */
int nmb_csv_write(const char* path, const char* header, const char* const* row_prefix, const void* body, int32_t body_is_f64,
                  int64_t n_rows, int32_t n_cols, int64_t ld, int32_t threads);

/* ---- ensemble ---------------------------------------------------------------------- */
int nmb_ensemble_create(NmbEnsemble** out, int32_t device, const NmbMember* members /*host*/,
                        int32_t n_members);
int nmb_ensemble_destroy(NmbEnsemble* ens);
int nmb_ensemble_size(const NmbEnsemble* ens, int32_t* n_members);
/* Which engine nmb_ensemble_train runs for these flags: 2 = pipelined tcgen05 kernel, 1 = generic tcgen05
 * engine, 0 = FP32 FFMA engine.  (Introspection for tests / benchmarks.) */
int nmb_ensemble_engine(const NmbEnsemble* ens, uint32_t flags, int32_t* engine);
/* steps already taken by each member (global_step of the train script :178), host out */
int nmb_ensemble_steps_done(NmbEnsemble* ens, int64_t* steps /*host [n_members]*/, void* stream);

enum {
  NMB_TRAIN_NO_ADAM = 1,     /* forward + loss + backward only (parity of gradients) */
  NMB_TRAIN_WRITE_GRADS = 2, /* store d(total)/d(param) into NmbMember.grads */
  NMB_TRAIN_KEEP_ACTS = 4,   /* keep x_recon in scratch instead of overwriting it with its gradient */
  NMB_TRAIN_FP32 = 8,        /* run every dense stage on the FP32 FFMA engine (bit-stable trajectories) instead of
                                the default tcgen05 engine (error-compensated BF16x3 products, FP32 accumulate) */
  NMB_TRAIN_TC_SIMPLE = 16,  /* tcgen05 engine without the operand pipeline (the generic engine that also serves
                                architectures the pipelined kernel does not cover, e.g. hidden width > 127) */
  NMB_TRAIN_LOSS4 = 64,      /* loss_out rows hold 4 values: (total, kl, ll, head loss) -- losses['regression'] of
                                cVAE_multimodal_regression (cVAE.py:2343-2345); 0 for members without a head */
  NMB_TRAIN_LOSS8 = 128,     /* loss_out rows hold 8 values: (total, kl, ll = -(rec_health + rec_disease), cross-entropy,
                                rec_health, rec_disease, contrastive, 0) -- the loss dict of cVAE.py:2186-2195 */
  NMB_TRAIN_NO_STATS = 256,  /* do not update the BatchNorm running statistics (a repeated pass over the same minibatch) */
  NMB_TRAIN_RESIDENT = 32    /* pipelined engine: leave parameters and Adam moments in the kernel's lane-major master
                                layout after the call (no conversion back, none in at the next call).  NmbMember.params /
                                adam_m / adam_v are then STALE until nmb_ensemble_sync (logvar_out and the gPoE alphas
                                are always current); every libnmb call that reads them syncs by itself.  For runs that
                                call train repeatedly (per-epoch logging): saves two state conversions per call. */
};
/* The fused hot loop: for every member, n_steps minibatch steps of
 *   forward_multimodal -> loss_function_multimodal -> zero_grad -> backward -> optimizer1.step()
 * (multimodal_kfold_train_cvae_supervised.py:177-199, cVAE.py:1166-1196, torch Adam), one
 * launch for the whole ensemble.  Step s of a member uses rows [pos*batch, pos*batch+batch)
 * with pos = s mod ceil(n_rows/batch).
 *   eps_override : NULL (in-kernel Philox) or [n_members][n_steps][batch][latent] injected draws
 *   loss_out     : NULL or [n_members][n_steps][3] = (total, kl, ll) per step
 *                  (the values the train script prints at batch 0, :201-203) */
int nmb_ensemble_train(NmbEnsemble* ens, int64_t n_steps, const float* eps_override,
                       float* loss_out, uint32_t flags, void* stream);

/* Whole epochs for members with DIFFERENT steps-per-epoch (folds of unequal size, e.g. after the healthy-control
 * filter of multimodal_kfold_cvae_nmmlp.py:314): member i runs exactly n_epochs * ceil(n_rows_i / batch_i) steps --
 * the `for epoch ... for batch` nest of the train script :177-199 -- in the same single launch.
 *   loss_out : NULL or [n_members][n_epochs * max_i steps_per_epoch_i][3]; member i fills its first
 *              n_epochs * steps_per_epoch_i rows, the rest is left untouched. */
int nmb_ensemble_train_epochs(NmbEnsemble* ens, int64_t n_epochs, float* loss_out, uint32_t flags, void* stream);

/* After NMB_TRAIN_RESIDENT calls: bring NmbMember.params / adam_m / adam_v up to date (no-op otherwise). */
int nmb_ensemble_sync(NmbEnsemble* ens, void* stream);
/* The caller is about to WRITE NmbMember.params / adam_m / adam_v (e.g. load new weights): syncs, then drops the resident
 * copy so that the next train call reads the caller's buffers again. */
int nmb_ensemble_invalidate(NmbEnsemble* ens, void* stream);

/* Stand-alone Adam update over a packed buffer: optimizer1.step() (torch.optim.Adam defaults,
 * cVAE.py:1111-1116) for the per-step nn.Module API, where forward/backward and step() are
 * separate calls.  t = 1-based step count.  Entries with zero gradient and zero state do not move,
 * which reproduces torch's "skip parameters whose grad is None". */
int nmb_adam_step(float* params, const float* grads, float* adam_m, float* adam_v, int64_t n,
                  int64_t t, float lr, float beta1, float beta2, float adam_eps, void* stream);

/* Activations of the LAST executed step of one member (debug / per-step parity):
 * mu, logvar: fused latent [rows][latent]; x_recon[m]: [rows][input_dims[m]] (needs
 * NMB_TRAIN_KEEP_ACTS; 2 * n_mod entries for end-to-end members, health decoders first).  Any pointer may be NULL.
 * rows = size of that step's minibatch. */
int nmb_ensemble_peek(NmbEnsemble* ens, int32_t member, float* mu, float* logvar,
                      float* const* x_recon /*host array of device ptrs*/, int32_t* rows /*host*/,
                      void* stream);

/* Members with a supervised head: the head's output of the last executed step, [rows][4] floats per row --
 * regression: {fi_pred, -, -, -}; end-to-end: {logit 0, logit 1, -, -} (train-mode BatchNorm / dropout of that step). */
int nmb_ensemble_peek_head(NmbEnsemble* ens, int32_t member, float* out4, void* stream);

enum { NMB_RECON_MEAN = 0, NMB_RECON_SAMPLE = 1,
       NMB_RECON_GIVEN_Z = 2 /* decoders only: eps[i] holds z itself [n_rows[i]][latent] -- Decoder.forward /
                                cVAE.decode (cVAE.py:197-206, 426-428, 1135-1136); mu / logvar are not produced */,
       NMB_RECON_FP32 = 16 /* OR into `mode`: FP32 FFMA engine instead of the default tcgen05 (BF16x3) engine */,
       NMB_RECON_TC_SIMPLE = 32 /* OR into `mode`: generic tcgen05 engine instead of the pipelined forward-only program
                                   (the default wherever the pipelined training kernel covers the architecture) */,
       NMB_RECON_KEEP_PLANES = 64 /* OR into `mode`: the caller has not written NmbMember.params since the last
                                   nmb_ensemble_train* call, so the BF16 weight planes the training kernel left behind are
                                   current and are not rebuilt (honoured only if that call ran the pipelined engine and
                                   nothing invalidated the ensemble since) */ };
/* Test-time reconstruction for every member on its own rows:
 *   mode MEAN   : decode(mu)                       -- cVAE.pred_recon, cVAE.py:549-555
 *   mode SAMPLE : decode(mu + eps*exp(logvar/2))   -- cVAE_multimodal.pred_recon, cVAE.py:1198-1208
 * xc[i*NMB_MAX_MOD+m]   : packed rows of member i, modality m ([n_rows[i]][ldx])
 * xhat[i*NMB_MAX_MOD+m] : out [n_rows[i]][input_dims[m]]
 * mu/logvar[i]          : optional out [n_rows[i]][latent] (fused latent; pred_latent :540-547)
 * eps[i]                : optional injected draws [n_rows[i]][latent] for SAMPLE (else Philox stream 1)
 * All pointer tables and n_rows are HOST arrays. */
int nmb_ensemble_reconstruct(NmbEnsemble* ens, const float* const* xc, const int32_t* n_rows,
                             int32_t mode, const float* const* eps, float* const* xhat,
                             float* const* mu, float* const* logvar, void* stream);

/* Members with a supervised head (NMB_HEAD_REGRESSION): the head's prediction for every row of xc -- `fi_pred` of
 * cVAE_multimodal_regression.forward_multimodal (cVAE.py:2309-2332): encode, fuse, z = mu + eps*std (mode SAMPLE: the
 * reference samples at test time too, ..._regression.py:146-152; eps[i] injected or Philox), decode, residuals,
 * regressor.  out[i]: [n_rows[i]] (NULL entries / members without a head are skipped); xhat / mu / logvar: optional
 * tables as in nmb_ensemble_reconstruct (may be NULL).
 * NMB_HEAD_ENDTOEND members: out[i] = [n_rows[i]][2] logits of classifier(z) in EVAL mode (BatchNorm running statistics, no
 * dropout) -- cVAE_multimodal_endtoend.predict (cVAE.py:2198-2203) with mode MEAN; xhat has 2 * n_mod entries per member. */
int nmb_ensemble_head_predict(NmbEnsemble* ens, const float* const* xc, const int32_t* n_rows, int32_t mode,
                              const float* const* eps, float* const* xhat, float* const* mu, float* const* logvar,
                              float* const* out, void* stream);

/* The same for `n_sets` row sets per member in ONE launch (e.g. the training rows for the normative statistics and the
 * test rows of the test script, :83-113): entry (s, i) of every table sits at index s * n_members + i, for xc / xhat at
 * (s * n_members + i) * NMB_MAX_MOD + m.  Results are identical to n_sets separate calls; one work queue instead of
 * n_sets keeps the tail of the launch short. */
int nmb_ensemble_reconstruct_sets(NmbEnsemble* ens, int32_t n_sets, const float* const* xc, const int32_t* n_rows,
                                  int32_t mode, const float* const* eps, float* const* xhat, float* const* mu,
                                  float* const* logvar, void* stream);

/* ---- deviation scoring (streaming, HBM-bound) --------------------------------------- */
/* Batched over `n_seg` independent (member, modality) segments; tables are HOST arrays of
 * device pointers / sizes. x rows have stride ldx (packed rows) -- only the first d columns
 * are read; xhat rows have stride d. */

/* Per-ROI normative statistics over the reference rows (mask[i] != 0, or all rows if NULL):
 *   mean_d = mean_i r_id, std_d = population std_i r_id, r = (x - xhat)^2.
 * The ROI-space analogue of separate_latent_deviation (utils_vae.py:155-161); defined in
 * oracle/deviation.py.  out_stats[s] : [2][d] = (mean row, std row). */
int nmb_normative_stats(int32_t n_seg, const float* const* x, const int32_t* ldx,
                        const float* const* xhat, const uint8_t* const* mask,
                        const int32_t* n_rows, const int32_t* d, float* const* out_stats,
                        void* stream);

/* dev_roi = (x - xhat)^2 (test script :141), dev_subj = sum_d dev_roi / D (cVAE.py:1210-1211),
 * z = (dev_roi - mean_d) / std_d when stats[s] != NULL.  dev_roi / z may be NULL. */
int nmb_deviation(int32_t n_seg, const float* const* x, const int32_t* ldx,
                  const float* const* xhat, const float* const* stats, const int32_t* n_rows,
                  const int32_t* d, float* const* dev_roi, float* const* z,
                  float* const* dev_subj, void* stream);

/* Latent-space normative deviation, the reference's own z-score (utils_vae.py:155-161):
 *   z[i][k]  = (mu[i][k] - mean_t mu_train[t][k]) / sqrt(var_t mu_train[t][k] + exp(logvar[i][k]))   (np.var, ddof = 0)
 *   dev[i]   = sum_k |z[i][k]| / latent                                       (latent_deviation :155-157)
 * mu_train[s]: [n_train[s]][latent[s]] latent means of the reference (healthy-control training) rows; mu / logvar[s]:
 * [n_rows[s]][latent[s]] of the scored rows (pred_latent returns exp(logvar), cVAE.py:540-547).
 * out_z[s] ([n_rows][latent], separate_latent_deviation) and out_dev[s] ([n_rows]) may each be NULL. */
int nmb_latent_deviation(int32_t n_seg, const float* const* mu_train, const int32_t* n_train,
                         const float* const* mu, const float* const* logvar, const int32_t* n_rows,
                         const int32_t* latent, float* const* out_z, float* const* out_dev, void* stream);

/* ROC-AUC of every column of `scores[s]` ([n_rows][n_cols], row stride n_cols) against
 * binary labels[s] (uint8, 1 = patient): exact pair counting with tie half-credit
 * == sklearn roc_curve + auc (multimodal_kfold_cvae_group_analysis_1x1.py:123-124).
 * out_auc[s]: [n_cols] float64;  out_u2[s] (optional): [n_cols] uint64 pair counts
 * U2 = sum_{pos,neg} 2*[s_p > s_n] + [s_p == s_n]  (AUC = U2 / (2 n_pos n_neg)). */
int nmb_auc(int32_t n_seg, const float* const* scores, const uint8_t* const* labels,
            const int32_t* n_rows, const int32_t* n_cols, double* const* out_auc,
            unsigned long long* const* out_u2, void* stream);

/* The one exchange step of the multi-GPU split: one fixed-size float64 record per (member, modality) segment
 *   { subject AUC | per-ROI mean [d_max] | per-ROI std [d_max] | per-ROI AUC [d_max] | per-subject deviation [n_test_max] }
 * (zero / NaN padded) assembled from the flat result buffers of nmb_normative_stats ([2][d] per segment at o_stats[s]),
 * nmb_auc (per-ROI at o_auc[s], per-subject [n_seg]) and nmb_deviation (per-subject at o_subj[s]) -- the rows every rank
 * all-gathers before modalities are averaged (group analysis :212-215).  ALL pointers are DEVICE pointers.
 * out: [n_seg][1 + 3 d_max + n_test_max]. */
int nmb_member_records(int32_t n_seg, const float* stats, const int64_t* o_stats, const double* auc_roi, const int64_t* o_auc,
                       const double* auc_subj, const float* subj, const int64_t* o_subj, const int32_t* seg_d,
                       const int32_t* n_test, int32_t d_max, int32_t n_test_max, double* out, void* stream);

/* Mean of k score vectors (modality averaging, group analysis :212-215). */
int nmb_mean_rows(const float* const* src /*host table*/, int32_t k, int64_t n, float* out,
                  void* stream);

/* In-kernel eps stream exposed for tests: out[i] = eps(seed, step, element i, stream id). */
int nmb_philox_normal(uint64_t seed, uint64_t step, uint32_t stream_id, int64_t n, float* out,
                      void* stream);

/* Test hook for the tensor-core GEMM engine of the fused kernels: one CTA computes
 *   C[m][n] = sum_k A(m,k) * B(n,k),  A(i,k) = a[i*lda + k] (a_kmajor) or a[k*lda + i], same for B,
 * with the BF16x3 split on tcgen05.  lda/ldb multiples of 4 floats, 16-byte aligned bases.
 * Not part of the reference's interface. */
int nmb_debug_tc_gemm(const float* a, int32_t lda, int32_t a_kmajor, const float* b, int32_t ldb,
                      int32_t b_kmajor, float* c, int32_t ldc, int32_t m, int32_t n, int32_t k, void* stream);

/* Test hook: timeline trace of the pipelined training kernel.  buf (device, >= 8 * (3*items + 5*steps)
 * bytes, or NULL to switch tracing off) receives %globaltimer stamps of launch-local minibatch step
 * `step` of CTA 0: per epilogue item {arrive, start, end}, per MMA step {deps ready, tiles ready,
 * issued}, per producer step {deps ready, issued}.  tools/trace_tcp.py prints it. */
int nmb_debug_tcp_trace(uint64_t* buf, int32_t step);

#ifdef __cplusplus
}
#endif
#endif /* NMB_H_ */
