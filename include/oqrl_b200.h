/*
 * oqrl_b200 -- C ABI of the B200-native PQC fitness path.
 *
 * The reference (frederikbckl/OQRL) is pure Python and has no FFI; the seam the
 * path sits behind is three Python classes.  Each entry point below cites the
 * reference interface it replaces; oqrl_b200/{vqc,optim,dataset}.py are the
 * host-side mirrors that bind them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes, no torch / C++ types;
 *   - `d_*` = device pointer (caller-owned, e.g. a torch tensor's data_ptr()),
 *     `h_*` = host pointer; `stream` = a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Device entry points are asynchronous on
 *     `stream`; `*_host` entry points synchronise before returning;
 *   - return 0 on success, negative oqrl_status on error; the message is
 *     available from oqrl_last_error() (thread-local, never NULL);
 *   - no exceptions cross the boundary; handles are thread-compatible, not
 *     thread-safe (the reference never calls concurrently).  The library keeps
 *     per-device scratch (workspace, chunk-reduction buffers, staging): calls
 *     on one device must be stream-ordered with respect to each other -- issue
 *     them on one stream, or order the streams with events;
 *   - float32 arithmetic, run-to-run bit-deterministic (fixed-order reductions,
 *     no floating-point atomics) -- honours utils.py:19
 *     torch.use_deterministic_algorithms(True).
 */
#ifndef OQRL_B200_H
#define OQRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OQRL_ABI_VERSION 1

typedef enum oqrl_status {
    OQRL_OK = 0,
    OQRL_ERR_INVALID = -1,     /* bad argument / spec */
    OQRL_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    OQRL_ERR_UNSUPPORTED = -3, /* valid request outside the built kernels' range */
    OQRL_ERR_NOMEM = -4
} oqrl_status;

enum { OQRL_ENT_SEL_CNOT = 0, /* qml.StronglyEntanglingLayers default: CNOT(i -> i+r_l), r_l=(l mod (n-1))+1; agent.py:45 */
       OQRL_ENT_CZ_RING = 1 };/* CZ(i,(i+1) mod n) ring (BASELINE.json wording; not in the reference) */
enum { OQRL_FEAT_FIRST_K = 0, /* qml.AngleEmbedding: RY(x_i) on wire i, i < n_feat; agent.py:44 */
       OQRL_FEAT_TILED = 1 }; /* RY(x[i mod n_feat]) on every wire (builder-defined, configs 3-5) */

/* Circuit description.  Replaces the constructor arguments of
 * VQC(input_dim, output_dim, n_layers) -- agent.py:17-34 -- plus the template
 * choices hard-wired at agent.py:44-46.  Parameter layout is the reference's:
 * params[(l*n_qubits + wire)*3 + {0:phi,1:theta,2:omega}] (agent.py:39-43). */
typedef struct oqrl_spec {
    int32_t n_qubits;    /* wires; reference: input_dim = 4                    */
    int32_t n_layers;    /* StronglyEntanglingLayers depth; reference: 2       */
    int32_t n_out;       /* <Z_i> read out for i < n_out; reference: 2 actions */
    int32_t n_feat;      /* features per state; reference: 4                   */
    int32_t entangler;   /* OQRL_ENT_*                                         */
    int32_t reupload;    /* 0 = embed once (reference), 1 = before every layer */
    int32_t feature_map; /* OQRL_FEAT_*                                        */
    int32_t reserved;    /* 0; measurement aids: bit 0 disables the exact last-layer light-cone pruning, bit 1 runs the shared-memory class on its general kernel, bit 2 keeps the register class from staging the batch in shared memory */
} oqrl_spec;

typedef struct oqrl_dataset oqrl_dataset; /* opaque, HBM-resident transitions */

/* ---- library ------------------------------------------------------------ */
int oqrl_abi_version(void);
const char *oqrl_last_error(void);
/* SM count / compute capability of the current device (fails if it is not sm_100). */
int oqrl_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);

/* ---- dataset: replaces OfflineDataset.__init__ (dataset.py:8-43) ----------
 * One-time upload of the whole transition table to HBM.  Inputs are host
 * arrays already normalised by the Python mirror (actions squeezed to [N],
 * rewards f64->f32, terminals bool->f32).  Besides the raw columns the handle
 * keeps a half-angle table cos(x/2), sin(x/2) of every observation so the
 * fitness kernel never evaluates input trigonometry. */
int oqrl_dataset_upload(const float *h_obs, const float *h_next_obs, const int32_t *h_act,
                        const float *h_rew, const float *h_term, int64_t n_rows, int32_t n_feat,
                        void *stream, oqrl_dataset **out);
int oqrl_dataset_free(oqrl_dataset *ds);
int oqrl_dataset_shape(const oqrl_dataset *ds, int64_t *n_rows, int32_t *n_feat);
/* Row gather on the device -- the index-returning fast path that replaces the
 * list-of-tuples of OfflineDataset.sample (dataset.py:56-66).  Any output may be NULL. */
int oqrl_dataset_gather(const oqrl_dataset *ds, const int32_t *d_idx, int32_t n_idx,
                        float *d_obs, float *d_next_obs, int32_t *d_act, float *d_rew, float *d_term,
                        void *stream);

/* ---- forward: replaces VQC.forward (agent.py:48-61) -----------------------
 * q[p][b][i] = <Z_i> of the circuit with parameters d_params[p] on input row
 * d_x[b].  P = 1 serves policy_net(x) / act(); P > 1 evaluates a population. */
int oqrl_qvalues(const oqrl_spec *spec, const float *d_params, int32_t n_cand,
                 const float *d_x, int32_t n_rows, float *d_q, void *stream);

/* ---- fitness: replaces the list comprehension over
 * GAOptimizer._evaluate_fitness (optim.py:64-173, :204) ----------------------
 * fitness[p] = -mean_t (Q_p(s_t)[a_t] - (r_t + (1-d_t)*gamma*max_a Q_p(s'_t)[a]))^2,
 * both Q evaluated with candidate p's own weights (optim.py:152,162).
 * `d_idx` selects the generation's transitions from the resident dataset.  The reference's fancy indexing
 * raises IndexError on a row outside the table (dataset.py:60-66); device indices cannot be checked without a
 * host round trip, so every kernel family reads row 0 instead of out of bounds and writes NaN for the fitness
 * values it computed from such a row (oqrl_fitness_host and the Python mirror check host indices up front and
 * fail with OQRL_ERR_INVALID / IndexError).  The dataset's actions must address a read-out (0 <= a < n_out,
 * recorded at upload): otherwise OQRL_ERR_INVALID, where the reference's torch.gather raises (optim.py:158). */
int oqrl_fitness(const oqrl_spec *spec, const float *d_params, int32_t n_cand,
                 const oqrl_dataset *ds, const int32_t *d_idx, int32_t n_trans,
                 float gamma, float *d_fitness, void *stream);
/* Same, for a batch that is not in the resident dataset (list[Experience]
 * re-stacked by the Python mirror: optim.py:73-112). All pointers on device. */
int oqrl_fitness_raw(const oqrl_spec *spec, const float *d_params, int32_t n_cand,
                     const float *d_obs, const float *d_next_obs, const int32_t *d_act,
                     const float *d_rew, const float *d_term, int32_t n_trans,
                     float gamma, float *d_fitness, void *stream);
/* Host-buffer variant: copies params and indices to the device, runs
 * oqrl_fitness, copies fitness back and synchronises.  h_* should be pinned. */
int oqrl_fitness_host(const oqrl_spec *spec, const float *h_params, int32_t n_cand,
                      const oqrl_dataset *ds, const int32_t *h_idx, int32_t n_trans,
                      float gamma, float *h_fitness, void *stream);

/* ---- multi-GPU: fitness all-gather fused into the kernel ---------------------
 * Population-sharded evaluation: this rank evaluates `n_cand` candidates that sit at
 * [cand_offset, cand_offset + n_cand) of the whole population and writes each fitness value directly
 * into the full fitness vector of EVERY rank: d_fitness_peers[r] is rank r's [P_total] float buffer as
 * mapped into this process (NVLink peer / symmetric memory; entry `rank` may be the local buffer).
 * The stores are issued by the fitness kernel itself as results become ready, so the exchange overlaps
 * the compute; the caller only needs a cross-rank barrier before reading.  Replaces
 * oqrl_fitness + ncclAllGather. */
int oqrl_fitness_scatter(const oqrl_spec *spec, const float *d_params, int32_t n_cand,
                         const oqrl_dataset *ds, const int32_t *d_idx, int32_t n_trans, float gamma,
                         float *const *d_fitness_peers, int32_t n_peers, int32_t cand_offset, void *stream);
/* Cross-rank barrier closing a generation of oqrl_fitness_scatter (replaces the barrier a collective
 * library would provide).  d_flag_peers[r] = rank r's flag vector (n_peers + 1 uint32, zero-initialised,
 * symmetric / peer-mapped memory) as mapped into this process.  `epoch` must grow by one per call on
 * every rank.  Asynchronous on `stream`; launched so that its launch latency overlaps the tail of the
 * preceding kernel.  Slot n_peers of the caller's own vector is set to 1 if a peer did not arrive
 * within ~10 s. */
int oqrl_peer_barrier(uint32_t *const *d_flag_peers, int32_t n_peers, int32_t rank, uint32_t epoch, void *stream);
/* Host-buffer variant of the sharded evaluation (what oqrl_fitness_host is to oqrl_fitness): this rank's block of
 * parameters is read in place from pinned host memory (pageable memory is staged), the row indices are staged, the
 * kernel's peer stores and oqrl_peer_barrier close the generation, and -- only if h_fitness_all is not NULL -- the
 * whole exchanged vector (n_total values) is copied into it; the stream is synchronised before returning.  The rank
 * that runs selection on the host passes a buffer, the others pass NULL and move no result at all. */
int oqrl_fitness_scatter_host(const oqrl_spec *spec, const float *h_params, int32_t n_cand, const oqrl_dataset *ds,
                              const int32_t *h_idx, int32_t n_trans, float gamma, float *const *d_fitness_peers,
                              int32_t n_peers, int32_t cand_offset, uint32_t *const *d_flag_peers, int32_t rank,
                              uint32_t epoch, float *h_fitness_all, int32_t n_total, void *stream);

/* ---- GA breeding: replaces the per-child torch calls of
 * GAOptimizer.optimize (optim.py:207-232), _crossover (:175-189) and _mutate (:191-197) ------------
 * One launch builds the next population [n_pop][n_genes] from the current one:
 *   rows 0..n_elite-1           = current rows d_order[0..n_elite-1]            (elitism, optim.py:212)
 *   rows n_elite + 2j, + 2j + 1 = children of pair j: parents a = d_order[pair[2j]], b = d_order[pair[2j+1]],
 *       child1 = where(mask, a, b) + noise[j][0], child2 = where(mask, b, a) + noise[j][1]
 *       (float32 gene + float64 noise, rounded to float32 -- the arithmetic of optim.py:193-197),
 *   truncated to n_pop rows (optim.py:232).
 * d_order [n_order] = candidate indices by descending fitness; n_elite <= n_order and parent_pool <= n_order
 * are required (the reference indexes population[rank] and raises IndexError past its end, optim.py:220-222);
 * d_pair [n_pairs][2] int32 parent ranks in [0, parent_pool), d_mask [n_pairs][n_genes] uint8,
 * d_noise [n_pairs][2][n_genes] f64 are the reference's RNG draws (choice / random < crossover_rate / normal),
 * produced on the host in the reference's stream order so that trajectories stay bit-identical.
 * d_next must not alias d_pop. */
int oqrl_ga_breed(const float *d_pop, int32_t n_pop, int32_t n_genes, const int32_t *d_order, int32_t n_order,
                  int32_t n_elite, const int32_t *d_pair, int32_t parent_pool, const uint8_t *d_mask,
                  const double *d_noise, int32_t n_pairs, float *d_next, void *stream);
/* Selection on the device: d_order[r] = index of the candidate with the r-th largest fitness, r < k
 * (the entries of `np.argsort(fitness)[::-1]`, optim.py:207, that the GA uses).  Order = a stable ascending
 * sort read backwards: exact ties resolve to the HIGHER index and NaN ranks FIRST (above +inf), so a NaN
 * candidate becomes the elite exactly as in the reference.  numpy's default sort leaves the order of exact
 * ties unspecified (its AVX-512 and scalar builds disagree even on a handful of elements); the stable order is
 * the reproducible member of that set.  Lets a whole optimize() call run without a host round trip. */
int oqrl_ga_rank(const float *d_fitness, int32_t n, int32_t k, int32_t *d_order, void *stream);
/* oqrl_ga_rank on the full fitness vector of a population sharded with oqrl_fitness_scatter (optim.py:204-209 across
 * GPUs: "all-gather of the fitness vector feeding selection"), k <= 8.  The kernel closes the generation itself: it
 * performs the cross-rank flag exchange of oqrl_peer_barrier (same flag vectors, same epoch rule) before it reads
 * d_fitness, and it is launched so that its launch latency overlaps the tail of the fitness kernel -- so a sharded
 * generation is fitness kernel, this kernel, oqrl_ga_breed, with no barrier launch in between.  Every rank runs it on
 * its own copy of the vector and obtains the same order. */
int oqrl_ga_rank_peers(const float *d_fitness, int32_t n, int32_t k, int32_t *d_order, uint32_t *const *d_flag_peers,
                       int32_t n_peers, int32_t rank, uint32_t epoch, void *stream);
/* oqrl_ga_breed with the randomness generated in the kernel (counter-based, keyed by seed and generation):
 * two distinct parents out of the best `parent_pool` per pair, uniform crossover mask, N(0, sigma) mutation
 * (optim.py:175-197 distributions).  No RNG-stream parity with the reference -- use oqrl_ga_breed with
 * host-drawn randomness for that. */
int oqrl_ga_breed_random(const float *d_pop, int32_t n_pop, int32_t n_genes, const int32_t *d_order,
                         int32_t n_elite, int32_t parent_pool, float crossover_rate, float sigma,
                         uint64_t seed, uint32_t generation, float *d_next, void *stream);

/* A whole GAOptimizer.optimize() call (optim.py:199-239) enqueued by ONE library call, no host work between the
 * generations: per generation the fused fitness kernel (oqrl_fitness on `ds` / `d_idx`), selection (oqrl_ga_rank with
 * k = max(n_elite, parent_pool)) and breeding -- the same kernels in the same order as the three separate entry
 * points, so the populations are bit-identical to driving them from the host (which costs ~25 us of interpreter and
 * binding time per call: a generation of the reference's default population is launch-bound without this).
 * Breeding follows the reference's RNG draws when d_pair / d_mask / d_noise are given -- [n_generations] x the arrays
 * of oqrl_ga_breed, n_pairs per generation -- and oqrl_ga_breed_random(seed, generation0 + g) when they are NULL.
 * The population ping-pongs between d_pop_a (the initial one) and d_pop_b, both [n_pop][n_genes]; *final_in_b says
 * where the offspring of the last generation live.  d_fitness [n_pop] receives the LAST generation's fitness vector,
 * d_best [n_genes] that generation's best individual (optim.py:234-239), d_order is scratch for k int32. */
int oqrl_ga_run(const oqrl_spec *spec, float *d_pop_a, float *d_pop_b, int32_t n_pop, const oqrl_dataset *ds,
                const int32_t *d_idx, int32_t n_trans, float gamma, int32_t n_generations, int32_t n_elite,
                int32_t parent_pool, const int32_t *d_pair, const uint8_t *d_mask, const double *d_noise, int32_t n_pairs,
                float crossover_rate, float sigma, uint64_t seed, uint32_t generation0, float *d_fitness,
                int32_t *d_order, float *d_best, int32_t *final_in_b, void *stream);

/* ---- introspection -----------------------------------------------------------
 * (measurement hooks -- FMA microbenchmarks, work models, the HBM-class pass list -- live in a separate
 * library: include/oqrl_b200_probe.h) */
/* Final statevector of ONE circuit in logical order (wire 0 = MSB), complex64
 * interleaved re,im: d_state[2 * 2^n].  Test hook for the large-n kernels. */
int oqrl_statevector(const oqrl_spec *spec, const float *d_params, const float *d_x,
                     float *d_state, void *stream);
/* Which kernel family a spec dispatches to: 0 = register-resident (n<=4),
 * 1 = shared-memory-resident, 2 = HBM-resident.  Negative = error. */
int oqrl_kernel_class(const oqrl_spec *spec);
/* Name of the fitness kernel a spec dispatches to (the name ncu lists), NULL on an invalid spec. */
const char *oqrl_kernel_name(const oqrl_spec *spec);
/* Number of kernel launches issued by this library since load (bench.py gpu_launches). */
int64_t oqrl_launch_count(void);
#ifdef __cplusplus
}
#endif
#endif /* OQRL_B200_H */
