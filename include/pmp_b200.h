/* pmp_b200.h -- C ABI of the B200-native potential-diffusion planning hot path.
 *
 * The reference (devinluo27/potential-motion-plan-release) is pure Python and has no FFI of
 * its own; the seam it offers is Python class identity at diffuser.models.* plus pickled
 * configs.  Every entry point below names the reference function it
 * replaces (file:line under the reference root).  They are bound with ctypes by
 * `potential-motion-plan-release_b200/pmp_b200/_lib.py` and are what a maintainer of the
 * reference would call from diffuser/models/{temporal_cond,diffusion_pb}.py and
 * diffuser/guides/{rm2d,kuka}_colchk.py.
 *
 * Conventions
 *   - all functions return 0 on success, a negative pmp_status otherwise; the message of the
 *     last failure on the calling thread is returned by pmp_last_error();
 *   - no function throws, none synchronises the device except pmp_finalize() and where stated;
 *   - pointers named *_dev are borrowed device pointers (cudaMalloc'ed / torch tensors) that
 *     must stay valid until the work enqueued on `stream` has completed; `stream` is a
 *     cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - tensors are dense row-major float32 unless said otherwise; trajectories are
 *     [batch, horizon, dim] exactly like the reference ("b h t", temporal_cond.py:247);
 *   - one handle per (process, device, model); a handle is not thread safe;
 *   - there is NO CPU fallback: every call fails with PMP_ERR_CUDA when no sm_100 device is
 *     usable.
 */
#ifndef PMP_B200_H_
#define PMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMP_ABI_VERSION 1
#define PMP_MAX_LEVELS 8
#define PMP_MAX_COMPOSE 8

typedef enum {
    PMP_OK = 0,
    PMP_ERR_ARG = -1,      /* bad argument / unsupported configuration */
    PMP_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
    PMP_ERR_STATE = -3,    /* call order: weights missing, not finalized, ... */
    PMP_ERR_NOMEM = -4
} pmp_status;

typedef struct pmp_model pmp_model;

/* Architecture of the energy U-Net == constructor kwargs of
 * TemporalUnet_WCond (diffuser/models/temporal_cond.py:67-243) */
typedef struct {
    int32_t transition_dim;              /* D: 2 (Maze2D) / 7 (Kuka) / 14 (Dual Kuka) */
    int32_t dim;                         /* base width, 64 in every shipped config */
    int32_t n_mults;
    int32_t dim_mults[PMP_MAX_LEVELS];   /* (1,4,8) or (1,4,8,12) */
    int32_t down_times;                  /* network_config['down_times'], >= n_mults if absent */
    int32_t time_mlp_config;             /* 0 default, 3 = Linear-act-Linear (temporal_dd.py:37-43) */
    int32_t act;                         /* 0 SiLU (energy_mode), 1 Mish */
    int32_t wall_embed_dim;              /* 64 */
    /* ViT1D wall encoder (diffuser/models/vit_vanilla.py:87-230) */
    int32_t vit_token_dim;               /* patch_size: 2 (x,y) / 3 (x,y,z | cx,cy,mode) */
    int32_t vit_octaves;                 /* 0 = no sinusoidal positional encoding */
    float   vit_norm_factor;             /* 30 */
    int32_t vit_depth;                   /* 3 */
    int32_t vit_mlp_ratio;               /* 4 */
    int32_t vit_act;                     /* 0: SiLU/SiLU (tf_config mish=False), 1: Mish/GELU */
    int32_t n_timesteps;                 /* T of the diffusion (for the time-embedding table) */
    int32_t reserved[8];
} pmp_arch;

const char *pmp_last_error(void);
int pmp_abi_version(void);
/* number of usable sm_100 devices visible to the process (0 => every other call fails) */
int pmp_device_count(void);

/* ---- model lifetime: replaces TemporalUnet_WCond.__init__ + load_state_dict
 *      (temporal_cond.py:67-243, utils/training_pbdiff.py:194-205) */
int pmp_model_create(const pmp_arch *arch, int device, pmp_model **out);
void pmp_model_destroy(pmp_model *m);
/* key = state-dict name without the leading "model." (e.g. "downs.0.0.blocks.0.block.0.weight");
 * data: float32, dense, reference layout; on_device != 0 => device pointer. Copies. */
int pmp_model_set_weight(pmp_model *m, const char *key, const float *data,
                         const int64_t *shape, int ndim, int on_device);
/* re-layouts the weights for the kernels, builds the time-embedding tables; synchronises. */
int pmp_model_finalize(pmp_model *m);
/* rows of the CFG batch processed per pass through the U-Net (bounds the workspace). */
int pmp_model_set_row_chunk(pmp_model *m, int rows);
/* 0: fp32 CUDA-core implicit GEMM everywhere; 1: tcgen05 BF16x3 tensor-core path (single CTA per
 * tile, direct epilogue stores); 2: operand-reuse variant of 1; 3 (default): tcgen05 BF16x3 on CTA
 * pairs (cta_group::2, 256-column tiles) with the TMA-store epilogue (16 epilogue warps for the
 * GroupNorm-backward epilogues); 4: path 3 with 16 epilogue warps everywhere.  1-4 fall back per op to
 * the next simpler kernel where a shape does not fit; paths 1-3 without 16-warp epilogues give
 * bit-identical results, the 16-warp epilogues sum the GroupNorm statistics in a different order
 * (differences ~1e-7 relative). */
int pmp_model_set_gemm_path(pmp_model *m, int path);
size_t pmp_model_workspace_bytes(const pmp_model *m);

/* ---- wall encoder: ViT1D.forward (vit_vanilla.py:202-230), called at temporal_cond.py:262.
 * walls_dev [B, n_tokens*token_dim] (metres, un-normalised) -> wemb_dev [B, wall_embed_dim].
 * Time independent, so the sampler evaluates it once per plan instead of once per step. */
int pmp_wall_embed(pmp_model *m, const float *walls_dev, int B, int wall_len,
                   float *wemb_dev, void *stream);

/* ---- potential gradient: TemporalUnet_WCond.forward in energy mode
 * (temporal_cond.py:245-334): u = UNet(x, t, w); E = 0.5*sum(u^2); eps = dE/dx by a
 * hand-written backward pass (replaces torch.autograd.grad at temporal_cond.py:321).
 * x_dev [R,H,D]; t_dev [R] int64; wemb_dev [n_cond,64]; rows >= n_cond are the
 * unconditional half of the classifier-free-guidance pair (wall embedding zeroed,
 * temporal_cond.py:272-279).  eps_dev [R,H,D]; energy_dev [R] or NULL; u_dev [R,H,D] or NULL. */
int pmp_unet_eps(pmp_model *m, const float *x_dev, const int64_t *t_dev,
                 const float *wemb_dev, int n_cond, int R, int H,
                 float *eps_dev, float *energy_dev, float *u_dev, void *stream);

/* ---- fused sampler update: GaussianDiffusionPB.p_mean_variance/p_sample
 * (diffusion_pb.py:161-195,236-246), pred_x_tm1_luo (:197-232), ddim_p_sample (:559-612),
 * pred_x_tm1_ddim (:670-721), apply_conditioning (helpers.py:168-172) and the K-way
 * composition of PolicyCompose (guides/policies_compose.py:286-297) in ONE kernel:
 *   eps  = eps_u[base] + sum_k w[k]*(eps_c[k]-eps_u[k])         (K=1: eps_u + w*(eps_c-eps_u))
 *   x0   = clamp(coef[0]*x - coef[1]*eps, -1, 1)
 *   DDPM : x' = coef[2]*x0 + coef[3]*x + coef[4]*z
 *   DDIM : mo = (x - coef[2]*x0)/coef[3];  x' = coef[4]*x0 + coef[5]*mo (+ coef[6]*z)
 *   x'[:,0,:] = start, x'[:,H-1,:] = goal                      (when start/goal != NULL)
 * coef: 8 host floats computed by the caller with the reference's own float32 ops, so the
 * update is bit-identical to PyTorch for identical eps.  z: noise_dev [B,H,D] or NULL, in
 * which case N(0,1) is drawn in-kernel from Philox4x32-10 (seed, offset).  x_out may alias x. */
typedef enum { PMP_STEP_DDPM = 0, PMP_STEP_DDIM = 1 } pmp_step_kind;
int pmp_step(int kind, const float *x_dev, const float *const *eps_c_dev,
             const float *const *eps_u_dev, const float *w, int K, int base,
             const float *coef, const float *noise_dev, uint64_t seed, uint64_t offset,
             const float *start_dev, const float *goal_dev, float *x_out_dev,
             int B, int H, int D, void *stream);
/* x = var_temp * N(0,1) then conditioning (diffusion_pb.py:256-258) */
int pmp_init_noise(float *x_dev, const float *noise_dev, uint64_t seed, uint64_t offset,
                   float var_temp, const float *start_dev, const float *goal_dev,
                   int B, int H, int D, void *stream);

/* ---- whole reverse-diffusion loop: GaussianDiffusionPB.p_sample_loop (:249-281),
 * ddim_p_sample_loop (:628-664) and PolicyCompose.ddpm/ddim_sample_loop
 * (policies_compose.py:270-350) for K composed models.  No host synchronisation inside. */
typedef struct {
    int32_t K;                              /* number of composed potentials (1 = plain CFG) */
    pmp_model *models[PMP_MAX_COMPOSE];
    const float *wemb_dev[PMP_MAX_COMPOSE]; /* [B,64] per model (pmp_wall_embed output) */
    float w[PMP_MAX_COMPOSE];               /* guidance weights */
    int32_t base;                           /* uncond_base_idx */
    int32_t B, H, D;
    int32_t kind;                           /* pmp_step_kind */
    int32_t n_steps;                        /* number of reverse steps to run */
    const int32_t *t_host;                  /* [n_steps] timestep fed to the U-Net at each step */
    const float *coef_host;                 /* [n_steps][8] per-step coefficients (see pmp_step) */
    const float *noise_dev;                 /* [n_steps+1][B,H,D] injected noise or NULL (Philox) */
    uint64_t seed;                          /* Philox seed when noise_dev == NULL */
    uint64_t row_offset;                    /* global index of row 0 (multi-GPU sharding) */
    float var_temp;
    const float *start_dev, *goal_dev;      /* [B,D] normalised */
    float *x_dev;                           /* [B,H,D] out: x_0 */
    float *trace_dev;                       /* [B,n_steps+1,H,D] or NULL (return_diffusion) */
} pmp_sample_args;
int pmp_sample(const pmp_sample_args *a, void *stream);


/* ---- training step: GaussianDiffusionPB.loss + loss.backward() (diffuser/models/diffusion_pb.py:480-545,
 * utils/training_pbdiff.py:118-131), i.e. the energy loss 0.5*mean(w*(dE/dx - noise)^2) and its gradient with respect to
 * every parameter (a second-order pass through the U-Net: forward, backward-data, tangent forward, reverse over the pair),
 * then torch.optim.Adam.step + EMA.update_model_average (training_pbdiff.py:133-140, train_utils.py:15-31).
 *
 * The caller owns ONE flat float32 parameter arena and ONE flat gradient arena (every nn.Parameter of the drop-in model is a
 * view into them, reference layouts).  pmp_train_bind names the arenas: keys[i] (state-dict name without "model.") starts at
 * element offsets[i]; every weight given to pmp_model_set_weight must appear. */
int pmp_train_bind(pmp_model *m, int n, const char *const *keys, const int64_t *offsets,
                   float *params_dev, float *grads_dev, int64_t total);
/* re-layouts the current parameter arena for the kernels (call after every optimizer step / load_state_dict);
 * rebuild_tables != 0 also rebuilds the sampler's time-embedding tables so pmp_sample sees the new weights. */
int pmp_train_sync_weights(pmp_model *m, int rebuild_tables, void *stream);
typedef struct {
    const float *x_noisy_dev;        /* [B,H,D] q_sample output with the start/goal rows conditioned (:483-487) */
    const int64_t *t_dev;            /* [B] diffusion step of every sample (:536) */
    const float *walls_dev;          /* [B,wall_len] */
    int32_t wall_len;
    const float *noise_dev;          /* [B,H,D] regression target (:505) */
    const float *keep_dev;           /* [B] 1 = keep the wall condition, 0 = concept dropped (temporal_cond.py:266-270) */
    const float *loss_weights_dev;   /* [H,D] loss_fn.weights (:112-134) */
    int32_t B, H;
    float *loss_dev;                 /* [1] out: the value `loss` returns (0.5 * weighted MSE) */
    float *eps_dev;                  /* [B,H,D] out or NULL: the model output dE/dx */
    float *energy_dev;               /* [1] out or NULL: sum of the batch's energies, the `engy_batch` of temporal_cond.py:324 */
} pmp_train_args;
/* grads arena += d loss / d parameters (zero it first, like optimizer.zero_grad()); no host synchronisation */
int pmp_train_loss_grad(pmp_model *m, const pmp_train_args *a, void *stream);
/* gradient buckets for an all-reduce that overlaps the reverse pass (DDP): bucket k covers arena elements
 * [lower_bounds[k], lower_bounds[k-1]) (lower_bounds descending, bucket 0 ends at the arena's end).  When the arena is
 * laid out in forward module order (embeddings / wall encoder, down path, middle, up path, final conv) the reverse pass
 * finishes it from the end towards the start, and pmp_train_loss_grad records an event per bucket as soon as every gradient
 * at or above its lower bound is final (otherwise all events fire at the end).  pmp_train_wait_bucket makes `stream`
 * (the communication stream) wait for bucket k of the most recent pmp_train_loss_grad. */
int pmp_train_set_buckets(pmp_model *m, int n, const int64_t *lower_bounds);
int pmp_train_wait_bucket(pmp_model *m, int k, void *stream);
/* q_sample + apply_conditioning (+ set_loss_noise) (:468-487, helpers.py:168-178): x_noisy = sqrt_acp[t] x0 +
 * sqrt_1m_acp[t] noise with rows 0 and H-1 copied from x0 (zero_cond_noise: their noise target zeroed in place) */
int pmp_q_sample(const float *x0_dev, float *noise_dev, const int64_t *t_dev, const float *sqrt_acp_dev,
                 const float *sqrt_1m_acp_dev, int B, int H, int D, int zero_cond_noise, float *x_noisy_dev, void *stream);
/* fused Adam (no weight decay / amsgrad; g is multiplied by grad_scale first, e.g. 1/world_size after an allreduce-sum)
 * + EMA of the parameters over flat arenas.  step counts from 1.  ema_mode 0: none, 1: ema = beta*ema + (1-beta)*p,
 * 2: ema = p (reset_parameters before step_start_ema). */
int pmp_adam_ema_step(float *p_dev, const float *g_dev, float *m_dev, float *v_dev, float *ema_dev, int64_t n,
                      float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                      int ema_mode, float ema_beta, void *stream);

/* ---- un-normalisation: LimitsNormalizer.unnormalize (datasets/normalization.py:151-162) */
int pmp_unnormalize(const float *x_dev, const float *lo_dev, const float *hi_dev,
                    float *out_dev, int64_t n_rows, int D, void *stream);

/* ---- Maze2D collision: RM2DCollisionChecker.check_1_traj_eps (guides/rm2d_colchk.py:76-93)
 * -> AbstractRobot._edge_fp (pb_diff_envs/robot/abstract_robot.py:127-153)
 * -> RectangleWallGroup.is_point_inside_wg (environment/rand_rec_group.py:216-237).
 * traj_dev [N,H,2] float32 un-normalised; env_id_dev [N] int32; cen_dev/hext_dev [E,nw,2]
 * float64; lo/hi: robot limits; legacy_div: 1 = numpy 1.23 semantics K=int(double(d)/eps),
 * 0 = numpy>=2 semantics K=int(d/float(eps)).  valid_dev [N] u8, nchk_dev [N] int32. */
int pmp_colchk_maze2d_swept(const float *traj_dev, const int32_t *env_id_dev, int N, int H,
                            const double *cen_dev, const double *hext_dev, int n_walls,
                            float lo_x, float lo_y, float hi_x, float hi_y,
                            double eps, double margin, int legacy_div,
                            uint8_t *valid_dev, int32_t *nchk_dev, void *stream);
/* per-waypoint scorer: check_single_traj / check_collision (guides/kuka_plan_utils.py:87-154).
 * ncol_dev [N] int32 (-1 if a waypoint is outside the limits, where the reference asserts),
 * ptcol_dev [N,H] u8 or NULL. */
int pmp_colchk_maze2d_points(const float *traj_dev, const int32_t *env_id_dev, int N, int H,
                             const double *cen_dev, const double *hext_dev, int n_walls,
                             float lo_x, float lo_y, float hi_x, float hi_y, double margin,
                             int32_t *ncol_dev, uint8_t *ptcol_dev, void *stream);

/* ---- Kuka / Dual-Kuka collision: KukaCollisionChecker.check_single_traj
 * (guides/kuka_colchk.py:81-104) with the pybullet contact query
 * (pb_diff_envs/robot/individual_robot.py:48-95, grouping.py:49-83) replaced by the analytic
 * sphere-link model defined in oracle/colchk.c (declared deviation: not pybullet meshes).
 * traj_dev [N,H,7*n_arms]; base [n_arms,3] host; sph [n_sph,5] host rows (link,cx,cy,cz,r);
 * boxc_dev/boxh_dev [E,nv,3] float32.  valid/nchk/ncol [N]; ptcol [N,H] or NULL. */
int pmp_colchk_kuka(const float *traj_dev, const int32_t *env_id_dev, int N, int H, int n_arms,
                    const float *base_host, const float *sph_host, int n_sph,
                    const float *boxc_dev, const float *boxh_dev, int n_vox, float table_z,
                    uint8_t *valid_dev, int32_t *nchk_dev, int32_t *ncol_dev,
                    uint8_t *ptcol_dev, void *stream);

/* ---- candidate selection: option_1_pick_1_traj (guides/rm2d_plan_env.py:137-235):
 * first valid candidate per problem else the last.  valid/nchk [n_prob*n_samp] ->
 * chosen [n_prob] int32, success [n_prob] u8, total_checks [n_prob] int32. */
int pmp_pick_first_valid(const uint8_t *valid_dev, const int32_t *nchk_dev, int n_prob, int n_samp,
                         int32_t *chosen_dev, uint8_t *success_dev, int32_t *total_dev, void *stream);

/* ---- introspection for tests / bench: number of kernels this library launched so far. */
uint64_t pmp_launch_count(void);
/* library-wide options.  "uncond_dedup" (default 1): compositions whose potentials are the SAME network
 * (config/comp: one model, K obstacle subsets) evaluate the wall-independent unconditional half once per step
 * instead of K times (bit-identical, policies_compose.py:286-297); 0 switches the sharing off. */
int pmp_set_option(const char *name, int value);
/* per-launch CUDA-event timing of the implicit-GEMM conv kernels (class 0: fp32 SIMT core,
 * class 1: tcgen05 core).  While enabled every conv launch is bracketed by two events on its
 * stream; pmp_profile_read waits for them, returns summed milliseconds, algorithmic FLOPs
 * (2*M*N*K of the convolution, independent of how many tensor passes the kernel makes) and
 * launch counts per class, and clears the records. */
int pmp_profile_enable(int on);
int pmp_profile_read(int n_classes, double *ms, double *flops, int64_t *launches);
/* debug (diagnostic builds only, `make diag`: the release library has the timers and every work-skipping probe
 * compiled out and always returns 1): cumulative clock64() role timers of conv_tc_kernel (env PMP_TC_TIMERS=1):
 * out8 = {mma_wait_tma, mma_wait_acc, mma_total, epi_wait_acc, epi_busy, tma_wait_stage, tiles, 0};
 * returns 1 when the timers are disabled */
int pmp_debug_counters(double *out8, int reset);
/* test hook: weight gradient of a stride-1 conv, dW[co][ci][k] += sum_{r,h} o[r,h,co] i[r,h+k-KS/2,ci] (two terms when o2 != NULL),
 * on the fp32 CUDA-core kernel (use_tc 0) or the tcgen05 kernel (use_tc 1; scratch of pmp_debug_wgrad_scratch_floats floats) */
int pmp_debug_wgrad(const float *o1, const float *i1, const float *o2, const float *i2, int H, int Co, int Ci, int KS, int R,
                    float *dw_dev, float *scratch_dev, int use_tc, int flags, void *stream);
size_t pmp_debug_wgrad_scratch_floats(int R, int H, int Co, int Ci, int KS);
/* debug: copy an internal activation of the last pmp_unet_eps call (first row chunk) to the
 * host; name e.g. "yhat:downs.0.0.blocks.0." ; returns element count or <0 */
int64_t pmp_debug_fetch(pmp_model *m, const char *name, float *host_out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* PMP_B200_H_ */
