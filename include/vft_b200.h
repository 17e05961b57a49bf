/*
 * vft_b200.h -- C-ABI of the B200-native hot path of VeryFastTree's neighbour-joining phase.
 *
 * This is the drop-in boundary.  The reference's only plug-in seam is the
 * per-element `Operations<Precision>` policy (src/operations/BasicOperations.h:16-39), whose
 * calls are 4..20 elements wide; a GPU cannot live behind that seam (the reference's own
 * CudaOperations.cu proves it).  The seam therefore moves one level up, to the loops of
 * NeighbourJoining.tcc that *call* those primitives.  Every entry point below replaces one
 * such loop and cites it; `veryfasttree_b200/csrc/B200Operations.h` is the C++ backend class a
 * maintainer registers next to AVX256Operations, and INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types cross this boundary;
 *   - every function returns 0 on success, a negative VFT_E* code otherwise
 *     (vft_last_error() gives the message; the C++ shim turns it into std::invalid_argument,
 *     which main.cpp:673-678 reports and exits 1);
 *   - buffers typed `void*` hold `float` when cfg.precision==32 and `double` when ==64
 *     (the reference's `numeric_t`);
 *   - node ids are the reference's: leaves 0..nSeqs-1, internal nodes nSeqs..2*nSeqs-1
 *     (NeighbourJoining.tcc:229-231); codes are 0..nCodes-1, VFT_NOCODE for gap/unknown
 *     (TransitionMatrix.h:7, NeighbourJoining.tcc:425,449-452);
 *   - all calls on one context are serialised by the caller (the batch *replaces* the
 *     reference's `omp for`); the library runs them on one CUDA stream per context.
 *
 * Limits (checked, never silently exceeded: VFT_EINVAL with a message from vft_last_error)
 *   - NJ phase: nPos <= 12 800 columns (shared-memory term buffer of the averageProfile kernel; checked by vft_nj_build
 *     up front, by vft_profile_average otherwise)
 *   - nSeqs < 2^30; K <= 4096 best hits per one-vs-all; merged candidate lists <= 4096 entries: with the reference's
 *     m = sqrt(N) top hits (K = 2m, lists <= 3m) that is N <~ 1.8 million taxa
 *   - vft_sh_support_batch: nPos <= 8 500 (the quartet's 3 x nPos site table in shared memory)
 *   - NJ driver: no topological constraints, no -slow, no 2nd-level top hits (-fastest)
 *   - vft_dist_*: one node, at most 16 ranks (one process per GPU); every rank must make the same sequence of calls
 *   - vft_ingest: rows travel in chunks of ~256 MB of text; nPos >= 1, no per-row terminators
 *
 * Arithmetic contract: results are bit-identical to the reference's "-mavx2" (no FMA)
 * build run with `-threads 1` -- same expression types (P vs double), same evaluation order,
 * same lane order in the P-typed dot products (AVX256Operations.tcc:5-26,58-138).
 */
#ifndef VFT_B200_H
#define VFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFT_NOCODE 127          /* TransitionMatrix.h:7 */
#define VFT_MAXCODES 20         /* DistanceMatrix.h:6   */

#define VFT_OK            0
#define VFT_EINVAL       -1     /* bad argument / precondition violated */
#define VFT_ENODEVICE    -2     /* no usable CUDA device (never falls back to the CPU) */
#define VFT_ECUDA        -3     /* CUDA runtime error */
#define VFT_ENOMEM       -4

/* lane order of the P-typed reductions (vector_multiply_sum / vector_multiply3_sum) */
#define VFT_REDUCE_SCALAR 0     /* BasicOperations.tcc:17-41 : left to right            */
#define VFT_REDUCE_AVX2   1     /* AVX256Operations.tcc:5-26 : lane accumulators + hadd */

typedef struct vft_ctx vft_ctx;

typedef struct vft_config {
    int64_t nSeqs;              /* unique sequences (NeighbourJoining ctor, NJ.tcc:220)      */
    int64_t nPos;               /* alignment columns                                          */
    int32_t nCodes;             /* 4 (nt) or 20 (aa), Options.h:43                            */
    int32_t precision;          /* 32 or 64 = sizeof(numeric_t)*8 (-double-precision)         */
    int32_t useMatrix;          /* Options.useMatrix: 0 => %different (nt default,
                                   VeryFastTree.cpp:96-98), 1 => tables from vft_upload_tables */
    int32_t reduction;          /* VFT_REDUCE_*                                               */
    int32_t device;             /* CUDA ordinal                                               */
    int32_t reserved;
    double  fPostTotalTolerance;/* Options.fPostTotalTolerance (Constants.h:37-38)            */
    int64_t nScratch;           /* extra dense profile rows, ids 2*nSeqs .. 2*nSeqs+nScratch-1: the
                                   temporaries of the ML phase (the reference's stack Profiles AB, CD, ...
                                   of MLQuartetOptimize NJ.tcc:1670-1745 and its upProfiles[], :3382-3434).
                                   Valid only as arguments of the likelihood entry points.  0 for NJ. */
} vft_config;

/* -- life cycle: replaces CudaOperations() ctor / configCuda (CudaOperations.cu:169-175) ---- */
int  vft_ctx_create(const vft_config *cfg, vft_ctx **out);
int  vft_ctx_destroy(vft_ctx *ctx);
const char *vft_last_error(void);
/* name of the implementation behind this ABI: "cuda-sm100a" for the product library,
   "oracle-cpu" for the test-only restatement under oracle/ */
const char *vft_backend_name(void);
/* Contexts come and go once per tree; the library parks the device / pinned blocks of a destroyed context
   in a process-wide cache (exact-size reuse, never the contents) so that the next vft_ctx_create does not
   pay the driver's allocation calls again.  This returns the parked blocks to the driver. */
int  vft_release_cached_memory(void);

/* -- model tables: after DistanceMatrix::setupDistanceMatrix (DistanceMatrix.tcc:102-153) -- */
/* distances[20][20], eigenval[20], eigentot[20], codeFreq[20][20] (row = code, first nCodes
   columns used; the reference pads rows to S, the padding is not part of the contract) */
int  vft_upload_tables(vft_ctx *ctx, const void *distances, const void *eigenval,
                       const void *eigentot, const void *codeFreq);

/* -- ingest: the character decoding of seqsToProfiles (NJ.tcc:415-457) + Uniquify (Alignment.cpp:494-526) ------
   text[nSeqs][nPos]: the alignment rows as the reference holds them after readAlignment (ASCII, no terminators);
   codesString: Options.codesString ("ACGT" for -nt, "ARNDCQEGHILKMFPSTWYV" otherwise; both cases decode, everything else
   -- '-' included -- is a gap, VFT_NOCODE).  The text is decoded and hashed on the device; rows are grouped in input order and
   every repeat is confirmed byte for byte, so the outputs are exactly the reference's: uniqueFirst[k] = the first row of the
   k-th distinct sequence (k < *nUnique), alnToUniq[i] = the distinct sequence row i belongs to, codes[*nUnique][nPos] = the
   decoded distinct rows in that order -- the `codes` argument of vft_upload_leaves / vft_nj_build.  All three arrays are
   caller-allocated for nSeqs rows. */
int  vft_ingest(const char *text, int64_t nSeqs, int64_t nPos, const char *codesString, int32_t device,
                int64_t *uniqueFirst, int64_t *alnToUniq, uint8_t *codes, int64_t *nUnique);

/* -- leaves: end of seqsToProfiles (NJ.tcc:382-534) ----------------------------------------- */
/* codes[nSeqs][nPos]; sets weights {0,1}, selfweight = nPos-nGaps (NJ.tcc:249-252), marks all
   leaves active, no internal nodes */
int  vft_upload_leaves(vft_ctx *ctx, const uint8_t *codes);

/* -- out-profile: outProfile (NJ.tcc:729-815) / updateOutProfile (NJ.tcc:943-1010) ---------- */
/* mean profile of the `n` nodes `ids` (ascending id order = the reference's accumulation order);
   ids==NULL => all currently active nodes.  Also recomputes codeDist (setCodeDist, :873-898). */
int  vft_outprofile_rebuild(vft_ctx *ctx, const int64_t *ids, int64_t n);
int  vft_outprofile_update(vft_ctx *ctx, int64_t old1, int64_t old2, int64_t newnode,
                           int64_t nActiveOld);

/* -- join: averageProfile (NJ.tcc:2067-2135) + the self-distance at NJ.tcc:3040-3043 -------- */
/* writes profile `out_id` = bionjWeight*id1 + (1-bionjWeight)*id2 (bionjWeight<0 => 0.5),
   records diameter[out_id], marks id1,id2 inactive (parent set, NJ.tcc:2905-2906) and out_id
   active, computes selfdist/selfweight[out_id] = profileDist(out,out). */
int  vft_profile_average(vft_ctx *ctx, int64_t out_id, int64_t id1, int64_t id2,
                         double bionjWeight, double diameter_out);
/* vft_profile_average followed by vft_outprofile_update(id1, id2, out_id, nActiveOld) in one launch:
   the join of NJ.tcc:3008 + :3035 when the out-profile is not rebuilt */
int  vft_profile_average_update(vft_ctx *ctx, int64_t out_id, int64_t id1, int64_t id2,
                                double bionjWeight, double diameter_out, int64_t nActiveOld);
int  vft_get_self(vft_ctx *ctx, int64_t id, double *selfdist, double *selfweight);
/* averageProfile (NJ.tcc:2067-2135, unweighted) for n INDEPENDENT items in one launch: profile out_id[k] = mean of id1[k], id2[k].
   Only the profile rows are written (no diameter, self distance, active flag): one tree level of recomputeProfiles. */
int  vft_profile_average_batch(vft_ctx *ctx, int64_t n, const int64_t *out_id, const int64_t *id1, const int64_t *id2);
/* recomputeProfiles (NJ.tcc:3474-3506) over a whole tree (arrays as in vft_nj_result): every 2-child internal node becomes the
   unweighted average of its children, children before parents -- as LEVEL-SYNCHRONOUS batches (one vft_profile_average_batch
   per height above the leaves; the reference's own threaded variant walks the same levels, :3483-3496).  This is the
   hand-over from the NJ slab to the ML phase: called after the tables were replaced by the transition matrix's
   (vft_upload_tables with transmat-as-distance-matrix, VeryFastTreeImpl.tcc:253-256) it re-expresses every internal
   profile in the new basis. */
int  vft_recompute_profiles(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child);

/* -- speculative join: the NEXT join's device work launched ahead of the host's decision ------------------------------
   The join loop is a dependency chain (search -> averageProfile -> distances of the new node -> bookkeeping -> search):
   the device idles while the host decides and the host idles while the device computes.  The caller's guess of the next
   join is right ~95 % of the time, so it may launch that join early, asynchronously, with NOTHING committed:
     launch : profile out_id (must be the next free id) = average of id1,id2 as vft_profile_average_update would build
              it; the updated out-profile goes to a shadow copy; then bare profileDist(out_id, pair_j[k]) (NJ.tcc:1167-1190,
              no diameter correction) and bare profileDist(out_ids[k], NEW out-profile) (out_ids may contain out_id).
              Raw values because the corrections (NJ.tcc:1120, :1046-1052) need diameter[out_id] / totdiam, which the
              host only knows once the join is decided; the caller finishes them in the same arithmetic.
     take   : the guess was right -- commits what vft_profile_average_update commits (per-node state, the out-profile:
              a pointer swap) and returns the results; self2 = {selfdist, selfweight} of the new node (numeric_t).
     discard: the guess was wrong -- nothing to undo.
   Any other entry point may be called between launch and take/discard; it sees the uncommitted state. */
int  vft_spec_join_launch(vft_ctx *ctx, int64_t out_id, int64_t id1, int64_t id2, double bionjWeight, int64_t nActiveOld,
                          const int64_t *pair_j, int64_t nPairs, const int64_t *out_ids, int64_t nOut);
int  vft_spec_join_take(vft_ctx *ctx, double diameter_out, void *pairDist, void *pairWeight, void *outDist,
                        void *outWeight, void *self2);
int  vft_spec_join_discard(vft_ctx *ctx);

/* -- out-distances: setOutDistance (NJ.tcc:1012-1053) --------------------------------------- */
/* fresh value for each ids[k] at this nActive/totdiam against the current out-profile.
   Pure: does not touch the context's own out-distance table (the caller decides what to commit,
   mirroring the reference's lazy refresh, NJ.tcc:1092-1098). */
int  vft_out_distance_batch(vft_ctx *ctx, const int64_t *ids, int64_t n, int64_t nActive,
                            double totdiam, void *outDist);
/* the loops NJ.tcc:257-260 / :4451-4464 / :3050-3055: every active node, committed to the
   context's table (consumed by vft_dist_one_vs_all).  outDist[maxnode]: entries of inactive
   nodes are left untouched. */
int  vft_out_distance_all(vft_ctx *ctx, int64_t nActive, double totdiam, void *outDist,
                          int64_t maxnode);

/* -- candidate lists: the distance half of setDistCriterion (NJ.tcc:1115-1122) as called from
      transferBestHits :4585-4612, uniqueBestHits :4823-4831, getBestFromTopHits :4287-4295 --- */
/* per pair: leaf x leaf -> seqDist (NJ.tcc:1601-1624); else profileDist (NJ.tcc:1167-1190)
   minus diameter[i]+diameter[j] (P arithmetic, :1120).  No criterion: the caller owns the lazy
   out-distance state that setCriterion needs. */
#define VFT_PAIRS_JOIN        0   /* setDistCriterion semantics as described above            */
#define VFT_PAIRS_PROFILE_RAW 1   /* bare profileDist (NJ.tcc:1167-1190): no leaf shortcut, no
                                     diameter correction -- the calls at NJ.tcc:3125-3127, :2945;
                                     j == -1 stands for the out-profile (BIONJ, NJ.tcc:2945-2946) */
int  vft_dist_pairs(vft_ctx *ctx, const int64_t *i, const int64_t *j, int64_t n, int32_t flags,
                    void *dist, void *weight);

/* the two list kinds above in ONE device call (one launch, one synchronisation): the per-join
   requests of the host loops are small, so their cost is latency, not bandwidth */
int  vft_eval_batch(vft_ctx *ctx, const int64_t *out_ids, int64_t nOut, int64_t nActive, double totdiam,
                    void *outDist, const int64_t *i, const int64_t *j, int64_t nPairs, int32_t flags,
                    void *dist, void *weight);

/* -- one-vs-all: setBestHit (NJ.tcc:3571-3639) + the psort/top-2m cut that always follows it
      (NJ.tcc:3930, :4471-4472) --------------------------------------------------------------- */
/* query vs every active node j < maxnode (self included, as the reference does), criterion
   from the committed out-distance table (precondition: vft_out_distance_all at this nActive, or
   the initial leaf state), then the K best in the order the reference's psort leaves them:
   criterion ascending, ties by j DESCENDING (oracle/psort_probe.cpp).  *nOut = min(K, nActive).
   Inactive nodes are not returned: their sentinel entries (NJ.tcc:3613-3617) sort after every
   active entry and can be appended by the caller. */
int  vft_dist_one_vs_all(vft_ctx *ctx, int64_t query, int64_t nActive, int64_t K,
                         int64_t *j_out, void *dist, void *weight, void *criterion,
                         int64_t *nOut);
/* the same sweep restricted to the candidate block jBegin <= j < jEnd: the unit of multi-GPU sharding
    -- each rank evaluates its block, the K records per rank are all-gathered and merged */
int  vft_dist_one_vs_all_range(vft_ctx *ctx, int64_t query, int64_t nActive, int64_t K, int64_t jBegin,
                               int64_t jEnd, int64_t *j_out, void *dist, void *weight, void *criterion,
                               int64_t *nOut);

/* -- one tree sharded over the GPUs of a node ----------------------------------------------------------------
   One process per GPU.  Every rank makes the SAME sequence of calls on a replicated profile slab (the join chain is
   deterministic, so no rank ever waits for another one's joins); the candidate axis of the three all-candidate sweeps is
   sharded and their results exchanged with one all-gather each:
     vft_dist_one_vs_all   rank r evaluates entries r, r+W, ... of the ascending active list, selects its K best; the W x K
                           fixed-size records are all-gathered and merged -- every rank returns the same K records
     vft_out_distance_all  the same strided share; the values are all-gathered into every rank's table
     vft_tophits_merge     the lists in W contiguous chunks; the saved lists are all-gathered
   Results are bit-identical to the unsharded calls (each distance is computed by exactly one rank with the same kernel).
   The reference has nothing comparable (CudaOperations.cu:8-12 maps OpenMP threads to devices); the contract is SURVEY 8e.
   vft_dist_init: collective over the `world` ranks; id128 = the 128-byte NCCL unique id made by vft_dist_unique_id on one
   rank and handed to the others by the launcher (torch.distributed / MPI / a file).  Contexts created on `device` afterwards
   are sharded.  Exchange: peer memory over NVLink (CUDA IPC + one push-and-wait kernel per exchange) when every GPU can map
   its peers, else ncclAllGather; VFT_EXCHANGE=nccl forces the latter.  vft_dist_init_host: the exchange goes through a
   caller-provided HOST all-gather instead (send: bytesPerRank bytes, recv: world x bytesPerRank, rank order; returns 0) --
   MPI / gloo bring-up, and what the CPU double of the tests uses. */
typedef int (*vft_allgather_fn)(const void *send, void *recv, int64_t bytesPerRank, void *user);
int  vft_dist_unique_id(void *id128);
int  vft_dist_init(int32_t rank, int32_t world, const void *id128, int32_t device);
int  vft_dist_init_host(int32_t rank, int32_t world, vft_allgather_fn fn, void *user, int32_t device);
int  vft_dist_finalize(void);
/* mode: 0 none (world 1), 1 NCCL, 2 peer memory, 3 host callback; counters since vft_dist_init */
int  vft_dist_info(int32_t *rank, int32_t *world, int32_t *mode, int64_t *nExchanges, int64_t *bytesExchanged);

/* -- top-hits refresh: the list-merging loop of topHitJoin's refresh branch (NJ.tcc:4477-4515) ------
   After a refresh's one-vs-all of `newnode` (NJ.tcc:4470-4472) the reference rebuilds the top-hit list
   of each of its m best hits iNode[l] from  (own list of iNode) + (the nAvail best hits of newnode):
     transferBestHits(.., updateDistances=false) :4580-4613  -- candidate j keeps its distance only when
                                                               iNode == newnode, otherwise it is unknown
     uniqueBestHits :4786-4833   -- psort by (i,j) (ties: reverse input order), first of each run of equal
                                    j kept, distance computed where unknown (setDistCriterion, :1115-1124),
                                    criterion for every survivor (setCriterion, :1085-1113)
     sortSaveBestHits :4535-4578 -- psort by criterion, the first m hits
   This entry point does those three steps for all nLists lists in one device pass.
   Inputs: own lists concatenated (ownOffset[nLists+1]); ownJ[k] = active ancestor of the stored hit
   (activeAncestor, :536-544; <0 if none), ownDist[k] = the cached distance, or any NEGATIVE value when the
   ancestor differs from the stored j (updateBestHit, :1626-1648) -- a negative distance is recomputed, as in
   the reference (:4826).  allJ/allDist[nAvail]: the sorted hits of newnode exactly as vft_dist_one_vs_all
   returned them (all active).  Criteria use the context's committed out-distance table: precondition
   vft_out_distance_all at this nActive (every out-distance fresh, so setCriterion has no side effect).
   Outputs: outCount[nLists] (= min(m, survivors)), outJ / outDist [nLists*m]. */
int  vft_tophits_merge(vft_ctx *ctx, int64_t newnode, int64_t nActive, int64_t m, int64_t nLists,
                       const int64_t *iNode, const int64_t *ownOffset, const int64_t *ownJ, const void *ownDist,
                       int64_t nAvail, const int64_t *allJ, const void *allDist,
                       int64_t *outCount, int64_t *outJ, void *outDist);

/* ============================================================================================
 * Likelihood kernels: pairLogLk and posteriorProfile under JC
 * (nucleotides, no transition matrix) or an eigen-decomposed rate matrix (GTR nt / JTT-WAG-LG aa),
 * with CAT rate categories.  Profiles are the same device slab; in the ML phase they hold the
 * rotated frequencies the reference keeps (VeryFastTreeImpl.tcc:253-256).
 * ==========================================================================================*/
/* after TransitionMatrix::create*() (TransitionMatrix.tcc:158-281).  NULL codeFreq => Jukes-Cantor
   (no transition matrix, NJ.tcc:1202).  codeFreq[(nCodes+1)][nCodes]: rows 0..nCodes-1 the codes,
   LAST row the NOCODE (gap) row transmat.codeFreq[NOCODE]; eigeninv[nCodes][nCodes] (row j = the
   vector of NJ.tcc:2426); eigeninvT[nCodes][nCodes] (nt only, NJ.tcc:2333; may be NULL for aa);
   eigenval[nCodes]; statinv[nCodes]. */
int  vft_upload_transmat(vft_ctx *ctx, const void *codeFreq, const void *eigenval, const void *eigeninv,
                         const void *eigeninvT, const void *statinv);
/* Rates (NJ.h:163-174) + the scalar options the two kernels read: rates[nRateCats] (P),
   ratecat[nPos], Options.MLMinRelBranchLength / MLMinBranchLength (Constants.h:32-36),
   Options.fastexp level 0..3 (BasicOperations.tcc:121-216) */
int  vft_sync_rates(vft_ctx *ctx, const void *rates, int64_t nRateCats, const int64_t *ratecat,
                    double MLMinRelBranchLength, double MLMinBranchLength, int32_t fastexpLevel);
/* pairLogLk (NJ.tcc:1192-1447) for n (pair, length) items -- the Brent abscissae of onedimenmin
   (NJ.tcc:7024), the per-node calls of treeLogLk (NJ.tcc:5123).  loglk[n] doubles.  siteLk: NULL,
   or [n][nPos] doubles receiving the per-site lkAB (what the reference multiplies into
   site_likelihoods[], :1263-1265; positions it skips get 1.0). */
int  vft_pair_loglk_batch(vft_ctx *ctx, const int64_t *i, const int64_t *j, const double *length,
                          int64_t n, double *loglk, double *siteLk);
/* posteriorProfile (NJ.tcc:2137-2447, exactML): profile out_id (an internal-node row) = posterior of
   the parent of id1,id2 at distances len1,len2 */
int  vft_posterior_profile(vft_ctx *ctx, int64_t out_id, int64_t id1, int64_t id2, double len1, double len2);

/* one tree LEVEL of recomputeMLProfiles (NJ.tcc:3508-3542) in one launch: n independent posteriorProfile items */
int  vft_posterior_profile_batch(vft_ctx *ctx, int64_t n, const int64_t *out_id, const int64_t *id1, const int64_t *id2,
                                 const double *len1, const double *len2);
/* the configuration the context was created with; *hasTransmat = a transition matrix is loaded (else Jukes-Cantor) */
int  vft_get_config(vft_ctx *ctx, vft_config *out, int32_t *hasTransmat);

/* -- whole-tree likelihood sweeps: recomputeMLProfiles (NJ.tcc:3508-3542) + treeLogLk
      (NJ.tcc:5114-5259) as LEVEL-SYNCHRONOUS batches over the entry points above (veryfasttree_b200/csrc/ml_host.cpp):
      one vft_posterior_profile_batch per tree level, one vft_pair_loglk_batch for all internal nodes, the terms
      added in the reference's post-order.  nChild[maxnode], child[maxnode*3], branchlength[maxnode] (numeric_t) as in
      vft_nj_result.  recomputeProfiles != 0: rebuild every 2-child internal profile bottom-up first.  The root's own
      profile row is used as scratch for the posterior of its first two children (NJ.tcc:5143-5146).
      leafCodes[nSeqs*nPos]: needed only for the Jukes-Cantor gap correction (NJ.tcc:5231-5257), may be NULL
      otherwise.  siteLoglk: NULL or [nPos]. */
int  vft_tree_loglk(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                    const void *branchlength, int32_t recomputeProfiles, const uint8_t *leafCodes, double *loglk,
                    double *siteLoglk);

/* setMLRates (NJ.tcc:5429-5488): the CAT approximation.  nRateCats candidate rates evenly spaced in log from 1/n to n
   (MLSiteRates, :5366-5377); for each, every site at that rate: recomputeMLProfiles + treeLogLk with per-site
   log-likelihoods (MLSiteLikelihoodsByRate, :5381-5408) -- nRateCats level-synchronous whole-tree sweeps; then the
   best category per site under the Gamma(3,1/3) prior, rates rescaled to average 1, profiles rebuilt.  Leaves the
   context with the chosen rates loaded (as vft_sync_rates).  rates[nRateCats] (numeric_t), ratecat[nPos];
   siteLoglk: NULL or [nRateCats*nPos].  The three scalars as in vft_sync_rates. */
int  vft_set_ml_rates(vft_ctx *ctx, int64_t root, int64_t maxnode, const int32_t *nChild, const int64_t *child,
                      const void *branchlength, int64_t nRateCats, double MLMinRelBranchLength, double MLMinBranchLength,
                      int32_t fastexpLevel, const uint8_t *leafCodes, void *rates, int64_t *ratecat, double *siteLoglk);

/* -- branch-length optimisation and ML NNI quartets as LOCK-STEP batches
      (veryfasttree_b200/csrc/ml_opt.cpp).  The reference optimises one branch at a time -- Brent's minimiser
      (onedimenmin / brent, NJ.tcc:7024-7190) asks for one pairLogLk per iteration -- and parallelises over OpenMP
      sections / tree partitions.  Here the n items of a call advance together: each round is ONE
      vft_posterior_profile_batch plus ONE vft_pair_loglk_batch carrying the pending request of every item.  Each
      item performs exactly the reference's `-threads 1` sequence of operations; the items of one call must be
      independent (no item reads a profile row another one writes).  Temporaries (the reference's stack Profiles)
      live in the scratch rows of vft_config.nScratch. ------------------------------------------------------------ */
typedef struct vft_ml_options {
    double  MLMinBranchLength;           /* Constants.h:30-31 (5e-4 float / 5e-9 double); must match vft_sync_rates */
    double  MLFTolBranchLength;          /* Constants.h:27-28 (0.001): Brent's fractional tolerance                 */
    double  MLMinBranchLengthTolerance;  /* Constants.h:24-25 (1e-4 / 1e-9): Brent's absolute tolerance             */
    double  closeLogLkLimit;             /* Constants.h:40 (5.0): star test and give-up thresholds of the NNI       */
    int32_t mlAccuracy;                  /* Options.h:62 (1): -mlacc; >= 2 always runs that many full rounds        */
    int32_t fastNNI;                     /* 1: star-topology test after the internal branch (NJ.tcc:1689-1698), as the
                                            reference's serial branch of MLQuartetNNI applies it (:4905-4918 -- at
                                            `-threads 1` it does so whatever its bFast argument says); 0: never, as
                                            in its OpenMP-sections branch (:4925-4951)                              */
} vft_ml_options;
typedef struct vft_ml_stats {            /* how the work of a call was batched */
    int64_t rounds;                      /* lock-step rounds                                 */
    int64_t loglkCalls, loglkItems;      /* vft_pair_loglk_batch calls / (pair,length) items */
    int64_t posteriorCalls, posteriorItems;
} vft_ml_stats;
void vft_ml_default_options(int32_t precision, vft_ml_options *opt);

/* MLPairOptimize (NJ.tcc:1790-1803) for n pairs: length[k] in = the starting guess (>= MLMinBranchLength), out = the
   optimum in [MLMinBranchLength, 6]; loglk[k] = pairLogLk there.  stats may be NULL. */
int  vft_ml_pair_optimize_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *idA, const int64_t *idB,
                                double *length, double *loglk, vft_ml_stats *stats);
/* MLQuartetNNI (NJ.tcc:4885-5004) for n quartets, each through MLQuartetOptimize (NJ.tcc:1650-1788) of up to three
   topologies per round: ids[4n] = profiles A,B,C,D (setupABCD order, NJ.tcc:1942-1974; D may be a scratch row holding
   an up-profile); len[5n] numeric_t in/out in the order A,B,C,D,internal; criteria[3n] (in: the caller's previous
   values, kept where a topology is not evaluated; out: log-likelihoods of AB|CD, AC|BD, AD|BC); choice[n] 0/1/2;
   starTest[n] (may be NULL) = 1 where the star test ended the item early (then only len[internal] is updated, as in
   the reference).  Uses scratch rows firstScratchRow .. firstScratchRow+3n-1.  No topological constraints. */
int  vft_ml_quartet_nni_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, void *len,
                              double *criteria, int32_t *choice, int32_t *starTest, int64_t firstScratchRow,
                              vft_ml_stats *stats);
/* The per-split body of testSplitsML (traverseTestSplitsML, NJ.tcc:6884-6952) for n independent splits: ids[4n] = A,B,C,D
   (setupABCD order), len[5n] numeric_t = the current lengths (already optimised for AB|CD).  loglk[3n]: MLQuartetLogLk
   (NJ.tcc:5410-5427) of AB|CD as it is, MLQuartetOptimize of AC|BD and AD|BC (the better one a second time when it is within
   closeLogLkLimit); siteLk[n][3][nPos]: the per-site likelihoods of each (the input of vft_sh_support_batch);
   choice[n]: the most likely topology; badSplit[n]: it beats AB|CD by more than treeLogLkDelta (0.1) -- the reference then
   reports support 0.  Scratch rows firstScratchRow .. +3n-1. */
int  vft_ml_split_test_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, const void *len,
                             double *loglk, double *siteLk, int32_t *choice, int32_t *badSplit, int64_t firstScratchRow,
                             vft_ml_stats *stats);
/* testSplitsML (NJ.tcc:6800-7000) over a whole tree (arrays as in vft_tree_loglk; node profiles current): the SH-like local
   support of every internal split.  The tests do not change the tree, so they are all independent: the up-profiles of the
   whole tree are built first (what getUpProfile builds lazily), then every split goes through vft_ml_split_test_batch's
   body in lock-step chunks and vft_sh_support_batch.  col[nBootstrap*nPos] from the caller's RNG (resampleColumns,
   NJ.tcc:705-716); support[maxnode] numeric_t: -1 where the reference sets none (leaves, root), 0 for a bad split;
   *nBadSplits as SplitCount.nBadSplits.  Needs one scratch row per internal node + 3 per split of a chunk. */
int  vft_ml_test_splits(vft_ctx *ctx, const vft_ml_options *opt, int64_t root, int64_t maxnode, const int32_t *nChild,
                        const int64_t *child, const void *branchlength, int64_t nBootstrap, const int64_t *col, void *support,
                        int64_t *nBadSplits, vft_ml_stats *stats);
/* chooseNNI (NJ.tcc:4836-4852), the minimum-evolution counterpart: for n quartets ids[4n] = A,B,C,D (node ids), the six
   profile distances of each (correctedPairDistances, NJ.tcc:1460-1488: bare profileDist, Options.pseudoWeight prior,
   logCorrect NJ.tcc:322-330 when logdist) evaluated as ONE vft_dist_pairs batch; criteria[3n] = d(AB)+d(CD), d(AC)+d(BD),
   d(AD)+d(BC) (lower is better), choice[n] 0/1/2 with the reference's tie rules.  No topological constraints. */
int  vft_choose_nni_batch(vft_ctx *ctx, int64_t n, const int64_t *ids, double pseudoWeight, int32_t logdist,
                          double *criteria, int32_t *choice);
/* The per-node body of optimizeAllBranchLengths (NJ.tcc:5044-5058) for n nodes: ids[3n] = the three profiles that
   meet at the node (child, child, up-profile -- or the three children of the root), len[3n] numeric_t in/out; two
   sweeps, each branch optimised against the posterior of the other two.  Scratch rows firstScratchRow .. +n-1. */
int  vft_ml_star_optimize_batch(vft_ctx *ctx, const vft_ml_options *opt, int64_t n, const int64_t *ids, void *len,
                                int64_t firstScratchRow, vft_ml_stats *stats);
/* optimizeAllBranchLengths (NJ.tcc:5006-5112) over a whole tree (arrays as in vft_tree_loglk; branchlength in/out;
   the node profiles must be current, e.g. after vft_tree_loglk(recomputeProfiles=1), and are kept current).
   VFT_ML_SCHEDULE_REFERENCE: the reference's post-order at `-threads 1`, up-profiles built lazily down the path from
     the root (getUpProfile, NJ.tcc:3382-3434) -- identical branch lengths; needs depth+2 scratch rows.
   VFT_ML_SCHEDULE_LEVELS: all nodes of one tree level in lock-step (up-profiles of the whole tree first, top-down);
     a node sees the state left by the previous LEVEL instead of by the previously visited node, the same kind of
     difference the reference's own tree-partitioned OpenMP mode has (NJ.tcc:5086-5107); needs one scratch row per
     internal node plus at least one more (the more, the wider the batches). */
#define VFT_ML_SCHEDULE_REFERENCE 0
#define VFT_ML_SCHEDULE_LEVELS    1
int  vft_ml_optimize_branch_lengths(vft_ctx *ctx, const vft_ml_options *opt, int64_t root, int64_t maxnode,
                                    const int32_t *nChild, const int64_t *child, void *branchlength, int32_t schedule,
                                    vft_ml_stats *stats);

/* -- SH-like local supports: SHSupport (NJ.tcc:1126-1165) for n quartets against ONE set of resampled
      columns (resampleColumns, NJ.tcc:705-716; col[nBootstrap*nPos] comes from the caller's RNG).  Per quartet: loglk[3] of the
      three topologies and their per-site likelihoods siteLk[3][nPos] as MLQuartetLogLk leaves them (NJ.tcc:5410-5427; the
      siteLk output of vft_pair_loglk_batch multiplied up by the caller).  Every resample is one ordered gather-sum over the
      columns for each topology (3 000 independent chains per quartet at the default 1 000 resamples); the logarithms are taken
      on the host with the reference's libm, so the supports are bit-identical.  support[n] = fraction of resamples in which the
      best topology's margin falls below the observed one. */
int  vft_sh_support_batch(vft_ctx *ctx, int64_t n, int64_t nBootstrap, const int64_t *col, const double *loglk,
                          const double *siteLk, double *support);

/* -- introspection for tests: a node's dense profile (weights[nPos], codes[nPos],
      vectors[nPos*nCodes], zero where the reference stores no vector); id==-1 => out-profile -- */
int  vft_get_profile(vft_ctx *ctx, int64_t id, void *weights, uint8_t *codes, void *vectors);
/* the counterpart: a dense profile written into an internal-node or scratch row (a profile the caller
   built itself, e.g. an up-profile it holds on the host); same three arrays */
int  vft_put_profile(vft_ctx *ctx, int64_t id, const void *weights, const uint8_t *codes, const void *vectors);

/* running totals of work done through this context (feeds roofline accounting, Debug.h:12-15) */
typedef struct vft_counters {
    int64_t seqOps;          /* leaf x leaf distances                    */
    int64_t profileOps;      /* profile distances (incl. out-profile)    */
    int64_t outprofileOps;   /* of which vs the out-profile              */
    int64_t profileAvgOps;   /* averageProfile calls                     */
    int64_t launches;        /* kernels launched (0 for the CPU oracle)  */
    int64_t algoBytes;       /* algorithmic bytes touched by the distance kernels */
    int64_t h2dBytes;        /* bytes copied host->device through this context */
    int64_t d2hBytes;        /* bytes copied device->host through this context */
    /* filled only when cfg.reserved & VFT_CFG_PROFILE: device time per kernel class, CUDA events
       on the launching stream (ms), and the algorithmic bytes of that class */
    double  msDist;          /* k_dist_pairs + k_one_vs_all + k_out_distance (the distance sweep) */
    double  msSelect;        /* top-K sort/merge/gather                                             */
    double  msProfile;       /* k_average + k_outprofile_update + k_outprofile_rebuild             */
    int64_t distLaunches;
    int64_t distBytes;
    /* the same event timings per kernel (VFT_CFG_PROFILE only): ms and launch counts, indexed as VFT_KERNEL_NAMES */
    double  msKernel[12];
    int64_t nKernel[12];
    int64_t bytesKernel[12];  /* algorithmic bytes of the distance kernels, always counted */
} vft_counters;
#define VFT_KERNEL_NAMES "k_eval(inline list)", "k_eval(batch)", "k_one_vs_all", "k_out_distance_all", "k_topk_select", \
                         "k_merge_prep+finish", "k_average", "k_outprofile_update", "k_outprofile_rebuild", "k_pair_loglk", \
                         "k_posterior", "k_nj_step"
#define VFT_CFG_PROFILE 1    /* vft_config.reserved bit: time every kernel with CUDA events */
int  vft_get_counters(vft_ctx *ctx, vft_counters *out);

/* device-side stopwatch on the context's stream (CUDA events; wall clock in the CPU oracle):
   brackets a host-driven sequence of calls for bench.py */
int  vft_timer_start(vft_ctx *ctx);
int  vft_timer_stop(vft_ctx *ctx, double *milliseconds);


/* ============================================================================================
 * Caller level (SURVEY.md §8 "next": the code either side of the kernels).
 * vft_nj_build runs the whole metric phase -- NeighbourJoining ctor tail (NJ.tcc:237-260) +
 * fastNJ() with the top-hits heuristic (NJ.tcc:2796-3155, :3746-4833) -- as a host-side batch
 * producer over the entry points above (veryfasttree_b200/csrc/nj_host.cpp).  The join order,
 * top-hit lists and branch lengths are those of the reference run with `-threads 1`.
 * ==========================================================================================*/
typedef struct vft_nj_options {
    double  tophitsMult;        /* Options.h:20  (1.0)  */
    double  tophitsClose;       /* Options.h:22  (-1.0 => log2(N)/(log2(N)+2), NJ.tcc:3747-3755) */
    double  topvisibleMult;     /* Options.h:25  (1.5)  */
    double  tophitsRefresh;     /* Options.h:28  (0.8)  */
    double  staleOutLimit;      /* Options.h:36  (0.01) */
    double  fResetOutProfile;   /* Options.h:38  (0.02) */
    int32_t nResetOutProfile;   /* Options.h:40  (200)  */
    int32_t bionj;              /* Options.h:18  (0): BIONJ-weighted joins, NJ.tcc:2921-2966 */
    int32_t prefetch;           /* 1 = batch the lazily refreshed out-distances / pair distances
                                   ahead of the host loops (default); 0 = fetch one at a time
                                   (same results, used by the tests to prove the prefetch is only
                                   a hint) */
    int32_t hostThreads;        /* OpenMP threads for the per-list host work of the refreshes and the
                                   leaf transfers (the reference parallelises the same loops,
                                   NJ.tcc:4477, :3838); 0 = min(16, cores).  Results do not depend on it. */
    int32_t deviceLoop;         /* 1 = the join loop (NJ.tcc:2857-3100) runs on the device: top-hit lists, visible sets
                                   and the lazily refreshed out-distances live in HBM, one thread block replays the
                                   reference's decisions and the grid evaluates the distances (nj_loop_logic.h);
                                   0 = the host-driven loop of round 1 (one synchronous call per join).  Same tree. */
    int32_t reserved;
} vft_nj_options;

void vft_nj_default_options(vft_nj_options *opt);

typedef struct vft_nj_result {
    /* caller-allocated, maxnodes = 2*nSeqs entries each (NJ.h:294-299) */
    int64_t *parent;            /* -1 for the root                                  */
    int32_t *nChild;
    int64_t *child;             /* [maxnodes*3]                                     */
    void    *branchlength;      /* P[maxnodes]                                      */
    /* optional traces for parity checks (NULL to skip) */
    int64_t *joins;             /* [(nSeqs-3)*2]  (i,j) of every join, in order     */
    int64_t *leafTopHits;       /* [nSeqs*m]      top-hit list of every leaf after
                                   setAllLeafTopHits (NJ.tcc:3746-4124), -1 padded  */
    /* filled in */
    int64_t root, maxnode, m;
    int64_t nSeeds, nCloseUsed, nRefreshTopHits, nVisibleUpdate, nHillBetter;
    int64_t nOutPrefetchHit, nOutSingleFetch, nPairPrefetchHit, nPairSingleFetch, nDeviceCalls;
    int64_t nSpecHit, nSpecMiss;/* speculative joins (vft_spec_join_*) that were taken / dropped */
    double  secondsLeafTopHits, secondsJoins, secondsTotal;
    double  deviceMsResident;   /* vft_timer around ctor tail + fastNJ: leaves already in HBM  */
    double  secondsEndToEnd;    /* host clock around everything incl. context + upload + result */
    double  secondsInCalls;     /* host clock spent inside the kernel-level ABI calls              */
    double  secondsHost[8];     /* host-only time: 0 leaf top-hits, 1 first top-visible, 2 join search,
                                   3 join bookkeeping (incl. 4), 4 topHitJoin (incl. 5), 5 refreshes  */
    vft_counters counters;
} vft_nj_result;

/* codes[nSeqs][nPos] as for vft_upload_leaves; tables==NULL unless cfg->useMatrix, else the four
   arrays of vft_upload_tables in that order. */
int  vft_nj_build(const vft_config *cfg, const vft_nj_options *opt, const uint8_t *codes,
                  const void *const tables[4], vft_nj_result *res);

#ifdef __cplusplus
}
#endif
#endif /* VFT_B200_H */
