/*
 * genjax_b200 C-ABI -- the drop-in boundary of the particle-parallel hot path.
 *
 * The reference (genjax-community/genjax @ 80ef143) is pure Python and has no
 * FFI; its extension points are Python ABCs.  Each entry point below names the
 * reference interface whose *batched* (jax.vmap-ed) execution it replaces.
 * Paths are relative to /root/reference/src/genjax/_src/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated;
 *     nothing here allocates or frees caller memory;
 *   - calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*),
 *     they never synchronise;
 *   - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = argument
 *     validation failure (GJB_E_*);
 *   - particle arrays are row-major [n] or [n, d] float32 / int32;
 *   - RNG: Philox4x32-10, key = (key0, key1), counter =
 *     (idx_lo, idx_hi, chunk, site) with idx = idx_offset + local index (the
 *     GLOBAL particle / chain index), so results do not depend on how
 *     particles are sharded over GPUs.  Scalar sites share one block between
 *     the 4 particles of global quad idx >> 2 (slot idx & 3); vector sites
 *     and rejection samplers use one stream per particle (csrc/gjb_rng.cuh).
 *
 * Two libraries export these symbols:
 *   libgjb_core.so          : everything in section 1 (model independent)
 *   model_<hash>.so         : section 2, one library per captured @gen model
 */
#ifndef GENJAX_B200_H
#define GENJAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GJB_ABI_VERSION 16

#define GJB_E_ARG (-1)      /* null / misaligned pointer, negative size          */
#define GJB_E_RANGE (-2)    /* size beyond what the kernel supports              */
#define GJB_E_MODE (-3)     /* unknown mode / flag                               */

/* Every particle's integer mass is at most 2^36 and the total S is summed in uint64 and converted as a signed
 * 64-bit value: resampling / log-sum-exp entry points return GJB_E_RANGE for n_total >= 2^27 particles. */
#define GJB_MASS_MAX_PARTICLES (1LL << 27)

/* ------------------------------------------------------------------ 1. core */

int gjb_abi_version(void);

/* Device-side workspace sizes (bytes) for n particles. */
int64_t gjb_resample_workspace_bytes(int64_t n);

/*
 * max_i logw_i  ->  *wmax, a float in ordered-uint encoding so that unsigned
 * atomicMax orders it: b = bits(f); enc = (b >> 31) ? ~b : (b | 0x80000000);
 * -inf encodes as 0x007FFFFF, which is what gjb_wmax_reset stores.
 * Replaces the max pass of jax.scipy.special.logsumexp at
 * inference/smc.py:97,107.  `wmax` must have been reset with gjb_wmax_reset.
 */
int gjb_wmax_reset(uint32_t* wmax, void* stream);
int gjb_weight_max(const float* logw, int64_t n, uint32_t* wmax, void* stream);

/*
 * Exact integer mass of the weights relative to the max:
 *   tile_mass[b] = sum_{i in tile b} round(2^36 * exp(logw_i - M)),  tile = 2048
 * (uint64; integer addition is associative, so the total is independent of
 * the reduction tree and of the GPU count).  Second pass of logsumexp,
 * inference/smc.py:96-97.  `m_global` (device float*) overrides *wmax when
 * non-null (multi-GPU: the all-reduced max).
 */
int gjb_weight_mass(const float* logw, int64_t n, const uint32_t* wmax,
                    const float* m_global, uint64_t* tile_mass, void* stream);

/*
 * lse_out[0] = M (fp32 max as double), lse_out[1] = S (total mass as double,
 * exact below 2^53), lse_out[2] = log-mean-exp = M + log S - 36 log 2 - log n_total.
 * ParticleCollection.get_log_marginal_likelihood_estimate, inference/smc.py:96-97.
 */
int gjb_lse_finalize(const uint64_t* tile_mass, int64_t n, const uint32_t* wmax,
                     const float* m_global, int64_t n_total, double* lse_out,
                     void* stream);

/*
 * Systematic resampling over the exact integer CDF: ancestors[j] = i for
 * j in [cnt_{i-1}, cnt_i), cnt_i = clamp(ceil(C_i * n_total / S - u0), 0, n_total).
 * Replaces the user idiom `jax.random.categorical(key, logits)` per offspring
 * (docs/cookbook/inactive/inference/mapping_tutorial.ipynb cell 37; the
 * library's single draw is ParticleCollection.sample_particle,
 * inference/smc.py:102-109).
 */
typedef struct gjb_resample_args {
  const float* logw;          /* [n] local log-weights                            */
  int64_t n;                  /* local particles                                  */
  const uint32_t* wmax;       /* encoded max (gjb_weight_max / model kernel)      */
  const float* m_global;      /* nullable: overrides *wmax (multi-GPU global max) */
  const uint64_t* tile_mass;  /* from gjb_weight_mass                             */
  const uint64_t* c_offset;   /* nullable: mass of the shards before this one     */
  const uint64_t* s_total;    /* nullable: global mass (else local total)         */
  int64_t n_total;            /* global particle count                            */
  int64_t out_lo, out_n;      /* write offspring j in [out_lo, out_lo+out_n) to ancestors[j-out_lo] */
  int64_t anc_base;           /* global id of local particle 0                    */
  uint32_t key0, key1;        /* resample key words                               */
  uint64_t key_index;         /* resample key lane: u0 = u01(Philox(idx=key_index, site 0, chunk 0).x) */
  const uint32_t* key_dev;    /* nullable: {key0, key1, index_lo, index_hi} read on the device instead */
  int32_t* ancestors;         /* [out_n]                                          */
  double* lse_out;            /* nullable: {M, S, M + log S - 36 log 2 - log n_total} */
  uint32_t* wmax_next;        /* nullable: reset to -inf for the next step        */
  uint32_t* heavy_ws;         /* nullable scratch, GJB_HEAVY_WS_WORDS uint32: gjb_mass_resample_systematic parks the
                                 offspring ranges of particles that own whole 4096-slot windows here and, after a
                                 second grid barrier, ALL CTAs fill them (degenerate weights no longer serialise on
                                 the one CTA that owns the heavy particle)                                        */
} gjb_resample_args;

#define GJB_HEAVY_CAP 1024
#define GJB_HEAVY_WS_WORDS (4 + 3 * GJB_HEAVY_CAP)

int gjb_resample_systematic(const gjb_resample_args* a, void* stream);

/*
 * gjb_weight_mass + gjb_resample_systematic as ONE cooperative launch (a grid
 * barrier instead of a kernel boundary; the integer masses stay in registers
 * between the two phases).  a->tile_mass is scratch written by the kernel.
 * Needs one resident CTA per 2048-particle tile: gjb_mass_resample_fits(n) says
 * whether n qualifies on the current device (1) or the two-launch form must be
 * used (0).  Single-device form (no c_offset / s_total / m_global).
 */
int gjb_mass_resample_fits(int64_t n);
int gjb_mass_resample_systematic(const gjb_resample_args* a, void* stream);

/*
 * Multinomial resampling: offspring j draws r_j (64 bits of its Philox lane,
 * site 0 chunk 1), target = mulhi64(r_j, S), ancestor = upper_bound(C, target).
 * `cdf` is a caller workspace of n uint64 (filled here).
 */
int gjb_resample_multinomial(const float* logw, int64_t n, const uint32_t* wmax,
                             const uint64_t* tile_mass, uint64_t* cdf,
                             uint32_t key0, uint32_t key1, uint64_t idx_offset,
                             int64_t n_out, int32_t* ancestors, void* stream);

/* The same with {key0, key1} read on the device (the filter loop replays a CUDA graph with a new key table). */
int gjb_resample_multinomial_keydev(const float* logw, int64_t n, const uint32_t* wmax,
                                    const uint64_t* tile_mass, uint64_t* cdf, const uint32_t* key_dev,
                                    uint64_t idx_offset, int64_t n_out, int32_t* ancestors, void* stream);

/*
 * dst[j, :] = src[ancestors[j], :]  (row_bytes per row, multiple of 4).
 * ParticleCollection.get_particle / tree_map(lambda v: v[idx]),
 * inference/smc.py:90-91.
 */
int gjb_gather_rows(const void* src, const int32_t* ancestors, void* dst,
                    int64_t n_out, int32_t row_bytes, void* stream);

/*
 * The user-side MH accept step (tests/inference/test_requests.py:136-137, 190-191):
 *   check = log(uniform(key)) < w;  trace = tree_map(lambda new, old: where(check, new, old), new_trace, old_trace)
 * gjb_accept_mask: mask[i] = log(u[i]) < w[i];  gjb_select_rows: out[i, :] = mask[i] ? a[i, :] : b[i, :] (a / b may
 * be one broadcast row).
 */
int gjb_accept_mask(const float* u, const float* w, int64_t n, int32_t* mask, void* stream);
int gjb_select_rows(const int32_t* mask, const void* a, const void* b, void* out, int64_t n, int32_t row_bytes,
                    int32_t a_broadcast, int32_t b_broadcast, void* stream);

/* Effective sample size (sum w)^2 / sum w^2, w = exp(logw - lse3[0]) (lse3 from gjb_lse_finalize); out: 1 double. */
int gjb_weight_ess(const float* logw, int64_t n, const double* lse3, double* out, void* stream);

/* ----------------------------------------------- 1b. multi-GPU (one process per GPU)
 *
 * Global resampling across R ranks: particles are block
 * partitioned (rank r owns global ids [r*n_per_rank, (r+1)*n_per_rank)), the
 * CDF is the concatenation of the ranks' exact integer CDFs, and every rank
 * writes the ancestors of ITS parents' offspring straight into the owning
 * rank's ancestor buffer over NVLink (peer pointers from a symmetric-memory
 * rendezvous).  The three per-step exchanges (global max, rank masses, barrier)
 * are single-CTA kernels that push 16 bytes to every peer's pad and poll their
 * own -- no NCCL call on the data path.
 */
#define GJB_MAX_RANKS 16

typedef struct gjb_peers {
  int32_t world;             /* 0 or 1: single device, tables ignored           */
  int32_t rank;
  int64_t n_per_rank;        /* particles per rank (multiple of 4)              */
  const void* base[GJB_MAX_RANKS];   /* the same buffer on every rank (peer mapped) */
  uint32_t div_mul, div_shr; /* owner = row / n_per_rank without a division (rows < 2^31):
                                n_per_rank == 1 ? row : umulhi(row, div_mul) >> div_shr, with p = 31 + ceil(log2 n_per_rank),
                                div_mul = ceil(2^p / n_per_rank), div_shr = p - 32 (host: gjb_peers_set_divisor) */
} gjb_peers;

/* Fills div_mul / div_shr for p->n_per_rank (host helper, no GPU work). */
int gjb_peers_set_divisor(gjb_peers* p);

#define GJB_XCHG_MAX 0       /* in: wmax (encoded)        out: m_global = max over ranks        */
#define GJB_XCHG_MASS 1      /* in: tile_mass[n_tiles]    out: c_offset (ranks before), s_total */
#define GJB_XCHG_BARRIER 2   /* all ranks reached this point; their earlier peer writes are visible */

typedef struct gjb_xchg_args {
  int32_t rank, world;
  int32_t mode;              /* GJB_XCHG_*                                       */
  int32_t n_tiles;           /* MASS: entries of tile_mass                       */
  uint64_t* pads[GJB_MAX_RANKS];     /* every rank's pad: uint64 [GJB_PAD_SLOTS][GJB_MAX_RANKS][2]; entry = value[31:0] | tag<<32,
                                        value[63:32] | tag<<32 (the 32-bit tag rides inside each 8-byte store it validates) */
  const uint64_t* epoch;     /* device counter (bumped once per filter run)      */
  uint64_t tag_offset;       /* 1 <= tag_offset < 2^16; tag = (*epoch + 1) << 16 | tag_offset; pad slot tag_offset % GJB_PAD_SLOTS */
  const uint32_t* wmax;      /* MAX input                                        */
  const uint64_t* tile_mass; /* MASS input                                       */
  float* m_global;           /* MAX output                                       */
  uint64_t* c_offset;        /* MASS output                                      */
  uint64_t* s_total;         /* MASS output                                      */
} gjb_xchg_args;

int gjb_exchange(const gjb_xchg_args* a, void* stream);

/*
 * The same three hand-offs FUSED into the compute kernels (no extra launches):
 * the last CTA of a producer kernel pushes {value, tag} to every peer's pad,
 * every CTA of the consumer kernel polls its own pad before it needs the value.
 *   model kernel   : waits BARRIER(t-1) before gathering, pushes MAX(t)
 *   mass kernel    : waits MAX(t),  pushes MASS(t)
 *   resample kernel: waits MAX(t) + MASS(t), pushes BARRIER(t)
 * Exchange k of a run uses tag = (*epoch + 1) << 16 | k and pad slot k % 4 (k < 2^16).
 */
#define GJB_PAD_SLOTS 4
#define GJB_PAD_WORDS (GJB_PAD_SLOTS * GJB_MAX_RANKS * 2)   /* uint64 words per rank */

typedef struct gjb_link {      /* lives in DEVICE memory; constant for a plan     */
  int32_t rank, world;
  uint64_t* pads[GJB_MAX_RANKS];
  const uint64_t* epoch;
  uint32_t* counter;           /* zero-initialised ticket counter (self resetting) */
} gjb_link;

int gjb_weight_mass_linked(const float* logw, int64_t n, uint64_t* tile_mass, const gjb_link* link,
                           uint64_t wait_max, uint64_t push_mass, void* stream);

/*
 * PULL form of the global resample (no cross-rank stores, no barrier): the mass
 * kernel's last CTA also writes the inclusive prefix of this rank's tile masses
 * (tile_prefix, read by peers); every rank then resolves the ancestors of ITS
 * OWN offspring slots by scanning the parent tiles -- local or on a peer --
 * whose offspring range overlaps its slots.  Per step: 2 hand-offs (MAX, MASS);
 * log-weights and tile prefixes are double buffered across steps.
 */
int gjb_weight_mass_prefix_linked(const float* logw, int64_t n, uint64_t* tile_mass, uint64_t* tile_prefix,
                                  const gjb_link* link, uint64_t wait_max, uint64_t push_mass, void* stream);
int gjb_resample_systematic_pull(const gjb_resample_args* a, const gjb_peers* logw_peers,
                                 const gjb_peers* prefix_peers, const gjb_link* link, uint64_t wait_max,
                                 uint64_t wait_mass, void* stream);
int gjb_resample_systematic_linked(const gjb_resample_args* a, const gjb_peers* anc, const gjb_link* link,
                                   uint64_t wait_max, uint64_t wait_mass, uint64_t push_barrier, void* stream);
int gjb_epoch_bump(uint64_t* epoch, void* stream);

/* gjb_resample_systematic with every rank's ancestor buffer: offspring j is
 * written to anc.base[j / n_per_rank][j % n_per_rank].  a->ancestors is ignored,
 * a->out_lo must be 0 and a->out_n the global particle count. */
int gjb_resample_systematic_peers(const gjb_resample_args* a, const gjb_peers* anc, void* stream);

/* gjb_gather_rows whose source rows live on the rank that owns the (global) ancestor id. */
int gjb_gather_rows_peers(const gjb_peers* src, const int32_t* ancestors, void* dst, int64_t n_out,
                          int32_t row_bytes, void* stream);

/* ------------------------------------------- 1c. tile-exponent masses (single-launch filter step)
 *
 * The filter step that is ONE launch (section 2, gjb_model_pf_step) cannot spend a pass on the global maximum
 * before it forms the integer masses.  Each tile of GJB_TE_TILE consecutive particles (tile = global index / 2048:
 * a constant of the algorithm, not of the launch) takes its masses relative to its own power-of-two reference
 * 2^e, e = ceil(max_tile(logw * log2 e)); whoever consumes the tiles aligns them with exact right shifts:
 *     q_i = round(2^36 * 2^(logw_i * log2e - e)),  cdf[i] = inclusive prefix of q inside the tile,
 *     rec[p] = {mass = cdf of the tile's last particle, e}
 *     E = max_p e_p;  s_p = min(E - e_p, 63);  P_p = inclusive prefix of (mass_p >> s_p);  S = P_last
 *     C_i = P_{p-1} + (cdf[i] >> s_p)   (a monotone integer CDF: any CTA / GPU partition gives the same bits)
 *     offspring counts / ancestors exactly as gjb_resample_systematic; log-mean-exp = E ln 2 + log S - 36 ln 2 - log n_total.
 * CPU restatement: oracle/smc.py (te_tile_masses, te_cdf, resample_systematic_te).  Replaces, like section 1, the
 * logsumexp + categorical-per-offspring idiom (inference/smc.py:96-109; mapping_tutorial.ipynb cell 37).
 */
#define GJB_TE_TILE 2048
#define GJB_TE_MAX_TILES 4096         /* global tiles one resampling can span (8 388 608 particles) */
#define GJB_TE_E_NONE (-2147483647 - 1) /* exponent of a tile without a finite weight */

typedef struct gjb_tile_rec {
  uint64_t mass;             /* sum of the tile's masses, relative to 2^e       */
  int32_t e;                 /* the tile's reference exponent (GJB_TE_E_NONE: no finite weight) */
  int32_t reserved;
} gjb_tile_rec;

/* logw[n] -> cdf[ceil(n / 2048) * 2048] (padding repeats the last value), recs[ceil(n / 2048)]. */
int gjb_te_masses(const float* logw, int64_t n, uint64_t* cdf, gjb_tile_rec* recs, void* stream);

/*
 * What the LAST CTA of a filter-step launch leaves for the next launch (one per device and step parity): the global
 * exponent and mass, the inclusive prefix of the aligned tile masses of ALL ranks, and for every LOCAL window of 2048
 * offspring slots the range of parent tiles with offspring in it -- so that the consumers (512+ CTAs) do no prefix
 * work of their own.  On several GPUs every rank builds its own copy from the records all ranks mailed to it.
 */
typedef struct gjb_step_table {
  uint64_t S;                /* total aligned mass (0: no weight has mass)       */
  int32_t E;                 /* global exponent                                  */
  int32_t n_tiles_total;
  uint32_t tag;              /* the step's record tag, stored LAST (after a fence): consumers of the next launch spin on it
                                instead of waiting for a kernel boundary           */
  uint32_t reserved;
  uint64_t pre[GJB_TE_MAX_TILES];      /* inclusive prefix of (mass_p >> shf[p]), global tile order */
  int32_t win[GJB_TE_MAX_TILES][2];    /* local window w: first / last parent tile with offspring in it */
  uint8_t shf[GJB_TE_MAX_TILES];
} gjb_step_table;

/* Tile-record mailbox of one rank: uint64 [2 (step parity)][GJB_TE_MAX_TILES][GJB_TE_LL_WORDS]; record of global tile
 * p = {mass[31:0] | tag << 32, mass[63:32] | tag << 32, (uint32)e | tag << 32, tag << 32}: the 32-bit tag travels inside
 * each 8-byte store it validates (NCCL-LL style), so a reader needs no fence between flag and payload.  tag =
 * ((*epoch + 1) << 16 | step + 1) & 0xffffffff.  A record with the right tag also says that the tile's CDF row and
 * state rows are visible in the owner's L2 (the owner fences before mailing it). */
#define GJB_TE_LL_WORDS 4
#define GJB_TE_MAILBOX_WORDS (2 * GJB_TE_MAX_TILES * GJB_TE_LL_WORDS)

typedef struct gjb_step_link {   /* lives in DEVICE memory; constant for a plan */
  int32_t rank, world;
  int32_t tiles_per_rank;    /* ceil(n / 2048), the same on every rank          */
  int32_t reserved;
  uint64_t* mailbox[GJB_MAX_RANKS];  /* every rank's mailbox (peer mapped); [rank] is the local one */
  const uint64_t* epoch;     /* device counter, bumped once per filter run (gjb_epoch_bump) */
  uint32_t* ticket;          /* zero-initialised CTA ticket counter (self resetting) */
} gjb_step_link;

/*
 * The table of one step built by a launch of its own (one CTA of 1024 threads per device), running BESIDE the step
 * kernel: it polls this device's mailbox until the records of all tiles of all ranks carry the step's tag (that wait is
 * the cross-rank hand-off of the step), then writes the table.  With GJB_STEP_PDL it is launched with programmatic
 * stream serialization behind the step kernel, and the next step kernel behind it.
 */
typedef struct gjb_te_table_args {
  const gjb_step_link* link;
  int32_t step;
  uint32_t flags;            /* GJB_STEP_PDL                                      */
  int64_t slot_offset;       /* global slot of this device's particle 0           */
  int64_t n_local, n_total;
  const uint32_t* reskey;    /* {key0, key1, index_lo, index_hi} of the resampling of this step's weights */
  gjb_step_table* table_out;
  double* lse_out;           /* nullable: {E ln 2, S, log-mean-exp}               */
} gjb_te_table_args;

#define GJB_STEP_PDL 1u      /* launch with programmatic stream serialization: the launch may start while the previous
                                kernel on the stream drains; a step kernel draws its random numbers, then waits
                                (griddepcontrol.wait) before it touches anything the previous launch wrote */
#define GJB_STEP_FLAGWAIT 2u /* step kernel, table form: spin on the table's tag instead of waiting for the kernel boundary */
#define GJB_STEP_LIGHT 4u    /* step kernel with `link`: the last CTA leaves a RANK-level table only -- S, E and, in pre[0 .. world),
                                the inclusive prefix of the ranks' aligned masses (`reserved` = 1) -- and the consumers of the next
                                launch form the tile prefix of the one or two ranks their window's parents live on themselves, from
                                the records in their own mailbox (as the single-device table-free form does for its 512 tiles).  The
                                serial tail of a step shrinks from a 4096-tile prefix + window searches to one reduction. */
int gjb_te_table(const gjb_te_table_args* a, void* stream);

typedef struct gjb_te_resample_args {
  const uint64_t* cdf;       /* [n_tiles_local * 2048] this device's within-tile CDFs */
  const gjb_tile_rec* recs;  /* [n_tiles_total] tile records of ALL ranks, in global tile order (ignored with `table`) */
  const gjb_step_table* table; /* nullable: the table a filter-step launch left (then out_lo must be this device's
                                  slot offset, a multiple of 2048, and out_n its particle count) */
  const gjb_peers* cdf_peers; /* nullable DEVICE pointer: cdf of every rank (n_per_rank = particles per rank) */
  int32_t n_tiles_total;
  int32_t reserved;
  int64_t n_total;           /* global particle count (offspring slots)          */
  int64_t out_lo, out_n;     /* resolve offspring j in [out_lo, out_lo + out_n) into ancestors[j - out_lo] */
  const uint32_t* key_dev;   /* {key0, key1, index_lo, index_hi}: u0 = u01(Philox(idx = index, site 0, chunk 0).x) */
  int32_t* ancestors;        /* [out_n] global parent ids (identity when no weight has mass) */
  double* lse_out;           /* nullable: {E ln 2, S, log-mean-exp}              */
} gjb_te_resample_args;

int gjb_te_resample(const gjb_te_resample_args* a, void* stream);

/*
 * The per-step key table of a filter run, derived ON THE DEVICE from the run key's two words (threefry2x32-20, the host
 * key tree of core/key.py): row t = {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo, res_idx_hi, mn_k0, mn_k1} with
 * (k_prop, k_res) = split(fold_in(key, t)) -- `jax.random.fold_in / split` of the cookbook filter loop.  out: [T, 8].
 */
int gjb_pf_key_table(uint32_t key0, uint32_t key1, int32_t T, uint32_t* out, void* stream);

/* Raw Philox words / N(0,1) draws for RNG known-answer tests. */
int gjb_philox_fill(uint32_t key0, uint32_t key1, uint64_t idx_offset,
                    uint32_t site, uint32_t chunk, int64_t n, uint32_t* out4,
                    void* stream);
int gjb_normal_fill(uint32_t key0, uint32_t key1, uint64_t idx_offset,
                    uint32_t site, int64_t n, int32_t d, float* out,
                    void* stream);

/* -------------------------------------------------------- 2. per-model .so */

#define GJB_MAX_SITES 32   /* a Scan unrolled into its caller brings length x sites-per-step of them */
#define GJB_MAX_ARGS 16
#define GJB_MAX_RETS 16

/* site_flags bits */
#define GJB_SITE_SAMPLE 1u     /* draw the value (else read site_in)            */
#define GJB_SITE_WEIGHT 2u     /* add this site's logpdf to the weight          */
#define GJB_SITE_BCAST 4u      /* site_in holds ONE value shared by all particles */

/*
 * One fused launch over n particles of a captured static @gen model:
 * visits every site in program order; a site either samples (Philox lane of
 * the particle, site counter from 1 as static.py:260-263) or reads its value
 * from site_in; accumulates score = sum logpdf and weight = sum logpdf over
 * GJB_SITE_WEIGHT sites.  With the flags set by the host this single entry
 * point is the batched form of
 *   simulate  (generative_functions/static.py:254-278, 787-793)
 *   assess    (static.py:298-321, 983-989)
 *   generate / importance (static.py:341-380, 795-810;
 *              core/generative/generative_function.py:629-675)
 *   edit(Update)     (static.py:407-466; distributions/distribution.py:179-244)
 *   edit(Regenerate) (static.py:616-673; distribution.py:258-300)
 * weight_out[i] = (weight_in ? weight_in[i] : 0) + weight - (score_in ? score_in[i] : 0).
 */
typedef struct gjb_model_args {
  int64_t n;                 /* particles in this launch                        */
  uint64_t idx_offset;       /* global index of local particle 0                */
  uint32_t key0, key1;       /* batch key words                                 */
  const uint32_t* key_dev;   /* nullable: {key0, key1} read on the device instead (graph replay) */
  const int32_t* gather;     /* nullable: per-particle args are read at gather[i] */
  const gjb_peers* peer_args; /* nullable DEVICE array [GJB_MAX_ARGS]: per-particle arg i lives on rank gather[i] / n_per_rank */
  const gjb_link* link;      /* nullable: fused cross-rank hand-offs (multi-GPU filter)          */
  uint64_t wait_off;         /* != 0: poll BARRIER exchange wait_off before reading gathered rows */
  uint64_t push_off;         /* != 0: last CTA pushes *wmax as MAX exchange push_off              */
  const void* args[GJB_MAX_ARGS];      /* model args: per-particle arrays or shared blocks */
  float scalars[GJB_MAX_ARGS];         /* model args passed by value (host scalars)        */
  const void* site_in[GJB_MAX_SITES];  /* constrained / previous values        */
  void* site_out[GJB_MAX_SITES];       /* nullable: where to store site values */
  void* ret_out[GJB_MAX_RETS];         /* nullable: return-value leaves        */
  uint32_t site_flags[GJB_MAX_SITES];
  const float* score_in;     /* nullable: previous total score (update/regenerate) */
  const float* weight_in;    /* nullable: log-weights to accumulate onto        */
  float* score_out;          /* nullable                                        */
  float* weight_out;         /* nullable                                        */
  uint32_t* wmax;            /* nullable: atomic max of weight_out (ordered-uint) */
  /* Reference-maximum filter step (filter-flag instantiation only, idx_offset % 4 == 0): with a reference
   * *m_ref >= every weight of this launch known BEFORE the launch (an analytic bound of the incremental weight,
   * gen/bounds.py), the exact integer masses round(2^36 exp(w - *m_ref)) are accumulated per 2048-particle tile while
   * the weights are still in registers, so the step needs neither the max pass nor the mass pass:
   * gjb_resample_systematic(m_global = m_ref, tile_mass) follows directly.  tile_mass must be zero on entry;
   * tile_mass_clear[0 .. tile_mass_clear_n) (the buffer of the NEXT step) is zeroed by this launch.            */
  const float* m_ref;                  /* nullable */
  unsigned long long* tile_mass;       /* nullable; [ceil(n / 2048)] */
  unsigned long long* tile_mass_clear; /* nullable */
  int64_t tile_mass_clear_n;
  /* Single-pass filter step (with the reference-maximum fields above; one CTA per 2048 offspring slots, n <= 2048 *
   * GJB_PULL_MAX_TILES): before proposing, every CTA resolves the ancestors of ITS OWN slots from the PREVIOUS step's
   * weights -- output-slot ("pull") systematic resampling over the tile masses that step accumulated -- writes them to
   * pull_ancestors and gathers the previous state through them, so a step is ONE launch: resample(t-1) + gather +
   * propose + logpdf + masses(t).  pull_logw == NULL (first step): no resampling, args are read in place.          */
  const float* pull_logw;                  /* nullable: previous step's log-weights [n]                  */
  const unsigned long long* pull_tile_mass; /* previous step's tile masses (relative to *pull_m_ref)      */
  const float* pull_m_ref;                 /* previous step's reference maximum                          */
  const uint32_t* pull_key;                /* {key0, key1, index_lo, index_hi} of that resampling        */
  int32_t* pull_ancestors;                 /* [n] out: ancestors of this step's particles                */
  double* pull_lse;                        /* nullable out: {M, S, log-mean-exp} of the previous step    */
  int64_t pull_n_total;                    /* particle count the offspring counts are scaled to          */
} gjb_model_args;
#define GJB_PULL_MAX_TILES 2048

/* JSON description of the captured model (sites, args, layouts); static storage. */
const char* gjb_model_info(void);
int gjb_model_launch(const gjb_model_args* a, void* stream);

/*
 * The whole T-step bootstrap particle filter as ONE persistent cooperative
 * launch.  The reference has no filter class: this is the batched form of the
 * user idiom "vmap(step.importance) -> log_w += w -> categorical resample ->
 * gather" (docs/cookbook/inactive/inference/importance_sampling.ipynb cell 16,
 * mapping_tutorial.ipynb cell 37; reweight formula inference/smc.py:383).
 * The model's return leaves are the next state and feed back as its first
 * n_state (per-particle) arguments.  Per step t:
 *   A  gather x[anc] + propose + logpdf -> logw, running max      (grid barrier)
 *   B  exact integer mass of each CTA's weights relative to max   (grid barrier)
 *   C  CDF scan + systematic offspring ranges -> ancestors[t]     (grid barrier)
 * lse[t] = {M, S, log-mean-exp increment} as gjb_lse_finalize.
 */
typedef struct gjb_pf_args {
  int64_t n;                 /* particles on this device                        */
  int64_t n_total;           /* global particle count (== n on one device)      */
  uint64_t idx_offset;       /* global index of local particle 0 (multiple of 4) */
  int32_t T;                 /* filter steps                                    */
  int32_t record;            /* 1: state_buf / ancestors / logw keep all T steps */
  int32_t n_state;           /* state leaves = leading model args = return leaves */
  int32_t reserved;
  const uint32_t* keys;      /* [T, 8] {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo, res_idx_hi, 0, 0} */
  const void* state0[GJB_MAX_ARGS];     /* initial state leaves [n(, d)]        */
  void* state_buf[GJB_MAX_ARGS];        /* [slots, n(, d)]; slots = record ? T : 2 */
  int64_t state_stride[GJB_MAX_ARGS];   /* bytes between slots                  */
  const void* shared[GJB_MAX_ARGS];     /* shared args, indexed by model arg position */
  float scalars[GJB_MAX_ARGS];          /* scalar args, indexed by model arg position */
  const void* obs[GJB_MAX_SITES];       /* observed sites: [T, ...] values (null = proposed) */
  int64_t obs_stride[GJB_MAX_SITES];    /* bytes per step                       */
  uint32_t site_flags[GJB_MAX_SITES];
  float* logw;               /* [record ? T : 1, n] pre-resampling log-weights  */
  int32_t* ancestors;        /* [slots, n]                                      */
  double* lse;               /* [T, 3]                                          */
  uint32_t* wmax;            /* [2] scratch                                     */
  uint64_t* cta_mass;        /* [gjb_model_pf_grid(n)] scratch                  */
  uint32_t* barrier;         /* [2] scratch                                     */
} gjb_pf_args;

/* CTAs the persistent kernel uses for n particles (sizes cta_mass). */
int gjb_model_pf_grid(int64_t n);
int gjb_model_pf_run(const gjb_pf_args* a, void* stream);

/*
 * ONE launch per filter step (the default of inference/pf.py): every CTA owns GJB_TE_TILE offspring slots; it
 * resolves their ancestors from the PREVIOUS step's tile-exponent CDF (section 1c; systematic, output-slot form),
 * gathers the previous state through them, proposes, scores the observed sites, and publishes the within-tile CDF
 * and tile record of ITS OWN new weights -- resample(t-1) + gather + propose + logpdf + masses(t) with only
 * block-level synchronisation.  Batched form of the same user idiom as gjb_model_pf_run.  Filter-flag instantiation
 * only (observed sites weighted + broadcast, every other site sampled); the model's return leaves are the next
 * state.  prev_cdf == NULL (first step): no resampling, the state is read in place.
 */
typedef struct gjb_step_args {
  int64_t n;                 /* particles on this device                          */
  int64_t n_total;           /* particles the resampling spans (== n unless ranks resample globally) */
  uint64_t idx_offset;       /* RNG lane of local particle 0 (multiple of 4)      */
  int64_t slot_offset;       /* global offspring slot / parent id of local particle 0 (multiple of 2048; 0 on one device) */
  int32_t step;              /* t: tags this launch's tile records, selects the mailbox parity t & 1 */
  uint32_t flags;            /* GJB_STEP_*                                        */
  const uint32_t* key_dev;   /* row t of the key table: {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo, res_idx_hi, 0, 0};
                                the proposals use words 0-1, the last CTA words 2-5 (resampling of THIS step's weights) */
  const void* args[GJB_MAX_ARGS];      /* state leaves of the PREVIOUS step (pre-resampling) then shared blocks */
  float scalars[GJB_MAX_ARGS];
  const gjb_peers* peer_args; /* nullable DEVICE array [GJB_MAX_ARGS]: state leaf i of every rank */
  const void* site_in[GJB_MAX_SITES];  /* observed values of this step (one value shared by all particles) */
  void* state_out[GJB_MAX_RETS];       /* next state leaves [n(, d)]               */
  float* weight_out;         /* nullable: this step's incremental log-weights [n]  */
  const uint64_t* prev_cdf;  /* nullable (first step): previous step's within-tile CDFs (this device) */
  const gjb_peers* cdf_peers; /* nullable DEVICE pointer: prev_cdf of every rank   */
  const gjb_step_table* table_in;      /* the table the previous launch left (with prev_cdf), or NULL: */
  const gjb_tile_rec* prev_recs;       /* ... single device without table: the previous step's plain tile records; every
                                          CTA then forms the tile prefix itself (n_tiles_total of them)             */
  int32_t n_tiles_total;
  int32_t reserved;
  double* prev_lse;          /* nullable, with prev_recs: {E ln 2, S, log-mean-exp} of the PREVIOUS step (written by CTA 0) */
  const uint32_t* prev_key;  /* {key0, key1, index_lo, index_hi} of the previous step's resampling */
  int32_t* ancestors_out;    /* nullable [n]: the ancestors this launch resolved (previous step's) */
  uint64_t* cdf_out;         /* [ceil(n / 2048) * 2048] this step's within-tile CDFs */
  gjb_tile_rec* recs_out;    /* nullable: this step's plain tile records [ceil(n / 2048)] (the table-free form) */
  const gjb_step_link* link; /* nullable: mailboxes, epoch, ticket (the table form; required on several devices) */
  gjb_step_table* table_out; /* with link: built by this launch's last CTA; NULL: the CTAs only mail their records and
                                gjb_te_table builds the table beside this launch   */
  double* lse_out;           /* nullable: {E ln 2, S, log-mean-exp} of THIS step's weights (written by the last CTA) */
} gjb_step_args;

int gjb_model_pf_step(const gjb_step_args* a, void* stream);

/*
 * ALL T steps of the same filter in ONE cooperative launch (one device; every window of 2048 slots needs a co-resident
 * CTA: gjb_model_pf_steps_fits(n)): the per-step kernel's table-free body in a loop, ONE grid-wide barrier per step
 * where gjb_model_pf_step has a kernel boundary; the random numbers of step t + 1 are drawn before the barrier of step
 * t.  lse rows 0 .. T-2 and (record) ancestor rows 0 .. T-2 are written by this launch, the last step is resampled by
 * gjb_te_resample(cdf + ((T-1) & 1) * tiles * 2048, recs + ((T-1) & 1) * tiles).
 */
typedef struct gjb_steps_args {
  int64_t n;                 /* particles                                         */
  uint64_t idx_offset;       /* RNG lane of particle 0 (multiple of 4)            */
  int32_t T;                 /* filter steps                                      */
  int32_t record;            /* 1: state_buf / logw / ancestors keep all T steps  */
  const uint32_t* keys;      /* [T, 8] key table (gjb_pf_args.keys)               */
  const void* state0[GJB_MAX_RETS];     /* initial state leaves [n(, d)]           */
  void* state_buf[GJB_MAX_RETS];        /* [slots, n(, d)]; slots = record ? T : 2 */
  int64_t state_stride[GJB_MAX_RETS];   /* bytes between slots                     */
  const void* shared[GJB_MAX_ARGS];     /* shared args, indexed by model arg position */
  float scalars[GJB_MAX_ARGS];
  const void* obs[GJB_MAX_SITES];       /* observed sites: [T, ...] values (null = proposed) */
  int64_t obs_stride[GJB_MAX_SITES];    /* bytes per step                          */
  float* logw;               /* record ? [T, n] : [n] (the last step's weights)    */
  int32_t* ancestors;        /* record: [T, n] (row T-1 is left to gjb_te_resample); else unused */
  uint64_t* cdf;             /* [2, ceil(n / 2048) * 2048] scratch                 */
  gjb_tile_rec* recs;        /* [2, ceil(n / 2048)] scratch                        */
  double* lse;               /* [T, 3]                                            */
} gjb_steps_args;

int gjb_model_pf_steps_fits(int64_t n);
int gjb_model_pf_steps(const gjb_steps_args* a, void* stream);

/*
 * Batched MCMC drivers generated for the same model (one chain per lane).
 *   mh  : Rejuvenate-style random-walk proposal on the selected sites +
 *         accept `log(u) < alpha` (inference/requests/rejuvenate.py:70-94;
 *         tests/inference/test_requests.py:136-137,190-191)
 *   hmc : HMC.edit + accept (inference/requests/hmc.py:156-211); compat_stale_grad=1
 *         reproduces the carried-gradient quirk at hmc.py:186.
 * state: [n_chains, D] latent values (site order), logp: [n_chains].
 */
#define GJB_CHAIN_HAVE_LOGP 1u  /* logp[] holds the log-density of state[] on entry (MH)      */
#define GJB_CHAIN_NO_ACCEPT 2u  /* always move: the bare request.edit of the reference, whose
                                   weight (alpha_out) the caller feeds to its own accept step */

typedef struct gjb_chain_args {
  int64_t n;                 /* chains in this launch                           */
  uint64_t idx_offset;       /* global index of local chain 0 (RNG lane)        */
  uint32_t key0, key1;
  const void* args[GJB_MAX_ARGS];
  float scalars[GJB_MAX_ARGS];
  const void* site_in[GJB_MAX_SITES];  /* observed (unselected) site values, broadcast or per chain */
  uint32_t site_flags[GJB_MAX_SITES];
  float* state;              /* in/out [n, state_width] selected-site values (site order) */
  float* logp;               /* in/out [n] total log-density of the chain's trace */
  int32_t* accept_count;     /* in/out [n] (nullable)                           */
  float* alpha_out;          /* nullable [n]: weight of the LAST transition (w + bwd - fwd / HMC alpha) */
  int32_t state_width;       /* D: must equal the width the kernel was generated for */
  uint32_t flags;            /* GJB_CHAIN_*                                     */
  int32_t n_steps;           /* transitions per launch                          */
  int32_t step0;             /* global index of the first transition (RNG)      */
  float step_size;           /* MH proposal scale / HMC eps                     */
  int32_t n_leapfrog;        /* HMC L                                           */
  int32_t compat_stale_grad; /* reference-compat switch.  HMC: reproduce hmc.py:186 (the carried gradient); MH: take the
                                backward proposal arguments at the OLD state as rejuvenate.py:84-86 does */
} gjb_chain_args;

int gjb_model_mh_chain(const gjb_chain_args* a, void* stream);
int gjb_model_hmc_chain(const gjb_chain_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENJAX_B200_H */
