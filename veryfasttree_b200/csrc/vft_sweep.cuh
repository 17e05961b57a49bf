// vft_sweep.cuh -- the many-vs-one distance sweeps of the 20-state matrix mode (amino acids, BLOSUM45 / JTT eigenbasis):
//   MODE 0  setBestHit (NJ.tcc:3571-3646): one query against every active node        -> dist / weight / criterion / sort key
//   MODE 1  the all-node setOutDistance pass (NJ.tcc:257-260, :4451-4464): every node against the out-profile
//   MODE 2  the list merges of a top-hits refresh (NJ.tcc:4477-4515): m lists of candidates, each against its own query
// (included by vft_cuda.cu).
//
// What all three share: thousands of candidates are evaluated against ONE profile (the "query": a node, or the out-profile).
// profileDistPiece (NJ.tcc:900-916) of a candidate position with a KNOWN code c against the query is then a function of
// (position, c) only -- distances[cq][c] when the query's code is known, vector_multiply3_sum(fq, codeFreq[c], eigenval)
// otherwise -- i.e. exactly what the reference precomputes for its out-profile as `codeDist` (setCodeDist, NJ.tcc:873-898).
// k_query_tables builds that table for any query node ([Lp][20], same expressions, same bits), plus the dense vector the
// query presents to a candidate WITHOUT a code (its own vector or codeFreq[cq]) and its weights.  A candidate position then
// costs one byte (code) + one weight + ONE table gather; only the ~5 % of positions where the candidate itself carries a
// vector need the 20-wide 3-way product, and those are queued and evaluated lane-per-item so that they do not serialise the
// warp.
//
// Work decomposition: one CTA = 32 candidates ("rows"), SWP_NP producer warps + ONE consumer warp, chunks of 32 positions.
//   producer warp w, round r   chunk c = r*NP + w: stages the query's [32][20] table slice of the chunk in shared memory, then
//              8 units of 4 rows x 32 positions -- lane l handles row 4u + l/8, positions 4(l%8)..+3: one 32-bit code load,
//              one 128-bit weight load, 4 shared-memory gathers -> w1*w2 and the piece go into the chunk's [32][32] tile;
//              candidate positions with a vector are queued and evaluated lane-per-item, results patched into the tile;
//   consumer warp, round r+1   lane k adds row k of the NP tiles of round r, left to right, into ITS candidate's denom /
//              top: the reference's order, one dependent DADD per position and chain (NJ.tcc:1172-1183; top's term
//              w*piece is formed here in double).  Tiles are double-buffered; one __syncthreads per round.
// The ordered chains (1312 x 8.5 cycles per candidate) are the floor of a sweep; the producers run ahead of them, so a
// group of 32 candidates costs about its chain, and a sweep costs (groups / co-resident CTAs) chains.
// Bytes per candidate position: 1 (leaf) or 5 (internal node, fp32).
#pragma once

namespace {

template<typename P> struct QTab { const P *cd, *v, *w; size_t stride; };      // [nTab][Lp][20], [nTab][Lp][20], [nTab][Lp]; stride = Lp*20

// One table per query node (blockIdx.y); thread per position.
template<typename P>
__global__ void __launch_bounds__(128)
k_query_tables(Store<P> s, const int32_t *__restrict__ nodes, int64_t single, P *__restrict__ tcd, P *__restrict__ tv, P *__restrict__ tw) {
    const int64_t q = nodes != nullptr ? (int64_t) nodes[blockIdx.y] : single;
    const int64_t pos = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if (pos >= s.Lp) return;
    const size_t t = (size_t) blockIdx.y * (size_t) s.Lp;
    P *cd = tcd + (t + pos) * 20, *v = tv + (t + pos) * 20;
    P w = 0;
    uint32_t c = VFT_DEV_NOCODE;
    if (pos < s.L && q >= 0) {
        c = s.codes[q * s.Lp + pos];
        w = q < s.nSeqs ? (c != VFT_DEV_NOCODE ? (P) 1 : (P) 0) : s.weights[(q - s.nSeqs) * s.Lp + pos];
    }
    tw[t + pos] = w;
    P f[20], o[20];
    if (!(w > 0)) {
#pragma unroll
        for (int k = 0; k < 20; k++) { f[k] = 0; o[k] = 0; }
    } else if (c != VFT_DEV_NOCODE) {
#pragma unroll
        for (int k = 0; k < 20; k++) { f[k] = s.codeFreq[c * 20 + k]; o[k] = s.distances[c * 20 + k]; }      // NJ.tcc:905-907
    } else {
        load_vec<P, 20>(s.vecs + ((q - s.nSeqs) * s.Lp + pos) * 20, f);
        P e[20];
#pragma unroll
        for (int k = 0; k < 20; k++) e[k] = s.eigenval[k];
        for (int k = 0; k < 20; k++) {                                                                       // NJ.tcc:914-916
            P b[20];
#pragma unroll
            for (int i = 0; i < 20; i++) b[i] = s.codeFreq[k * 20 + i];
            o[k] = vec_mul3_sum<P, 20>(f, b, e, s.reduction);
        }
    }
    store_vec<P, 20>(cd, o);
    store_vec<P, 20>(v, f);
}

constexpr int SWP_QCAP = 256;                    // queued vector positions per producer warp (flushed above 128; a unit adds <= 128)
constexpr int SWP_NP = 3;                        // producer warps per CTA
template<typename P> struct SweepCfg {
    static constexpr int WS = sizeof(P) == 4 ? 36 : 34;        // tile row stride: 16-byte aligned rows, conflict-free 128-bit row reads
    static constexpr size_t tile = 2 * 32 * (size_t) WS * sizeof(P);   // w1*w2 and piece of one chunk
    static constexpr size_t qs = 32 * 21 * sizeof(P);          // one table slice, row stride 21 (conflict-free staging and gathers)
    static constexpr size_t perProducer = qs + SWP_QCAP * 2;
    static constexpr size_t bytes = 2 * SWP_NP * tile + SWP_NP * perProducer + 32 * 4;
    static constexpr int threads = (SWP_NP + 1) * 32;
};

template<typename P, int MODE>
__global__ void __launch_bounds__(SweepCfg<P>::threads, sizeof(P) == 4 ? 3 : 1)
k_sweep20(Store<P> s, QTab<P> qt, const int32_t *__restrict__ list, int stride, int offset,
          const int32_t *__restrict__ reqA, const int32_t *__restrict__ reqB, int cap,
          int64_t nSlots, int64_t query, int64_t nActive, double totdiam,
          P *__restrict__ dist, P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys, P *__restrict__ res) {
    typedef SweepCfg<P> Cfg;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int WS = Cfg::WS, NP = SWP_NP;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool consumer = wid == NP;
    P *tiles = reinterpret_cast<P *>(smemRaw);                                   // [2][NP] x { Wt[32][WS], Pc[32][WS] }
    unsigned char *mine = smemRaw + 2 * NP * Cfg::tile + (size_t) (consumer ? 0 : wid) * Cfg::perProducer;
    P *qs = reinterpret_cast<P *>(mine);                                          // [32][21]   (producers)
    uint16_t *queue = reinterpret_cast<uint16_t *>(mine + Cfg::qs);
    int32_t *ids = reinterpret_cast<int32_t *>(smemRaw + 2 * NP * Cfg::tile + NP * Cfg::perProducer);
    const uint32_t Lp = (uint32_t) s.Lp, nSeqs = (uint32_t) s.nSeqs;
    const int nChunks = (int) (Lp / 32), nRounds = (nChunks + NP - 1) / NP;
    const int64_t nGroups = (nSlots + 31) / 32;
    const int sub = lane >> 3, pl = (lane & 7) * 4;                      // my row within a unit, my first position within a chunk

    for (int64_t g = blockIdx.x; g < nGroups; g += gridDim.x) {
        const int64_t k = g * 32 + lane;
        int32_t id = -1;
        if (k < nSlots) {
            if (MODE == 2) { if (reqA[k] >= 0) id = reqB[k]; }
            else id = list[k * stride + offset];
        }
        if (__ballot_sync(full, id >= 0) == 0) continue;                  // (every warp of the CTA sees the same group: uniform)
        const size_t tab = MODE == 2 ? (size_t) ((g * 32) / cap) : 0;     // cap is a multiple of 32: one query per group
        const P *tcd = qt.cd + tab * qt.stride, *tv = qt.v + tab * qt.stride, *tw = qt.w + tab * (qt.stride / 20);
        if (wid == 0) ids[lane] = id;
        __syncthreads();
        double den = 0, top = 0;                                          // consumer: my candidate's chains
        for (int r = 0; r <= nRounds; r++) {
            if (!consumer) {
                const int ch = r * NP + wid;
                if (r < nRounds && ch < nChunks) {
                    // ---- producer: the tile of chunk ch ------------------------------------------------------------
                    P *Wt = tiles + ((size_t) (r & 1) * NP + wid) * (Cfg::tile / sizeof(P)), *Pc = Wt + 32 * WS;
                    const uint32_t p0c = (uint32_t) ch * 32;
                    // phase-1 loads first (they have the longest latency): codes + weights of my 4 positions in each unit
                    uint32_t c4[8];
                    P w4[8][4];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int32_t rid = __shfl_sync(full, id, u * 4 + sub);
                        c4[u] = 0x7F7F7F7Fu;
#pragma unroll
                        for (int i = 0; i < 4; i++) w4[u][i] = 0;
                        if (rid >= 0) {
                            c4[u] = *reinterpret_cast<const uint32_t *>(s.codes + (uint64_t) (uint32_t) rid * Lp + p0c + pl);
                            if ((uint32_t) rid >= nSeqs) load_vec<P, 4>(s.weights + (uint64_t) ((uint32_t) rid - nSeqs) * Lp + p0c + pl, w4[u]);
                        }
                    }
                    // the query's table slice: lane p owns position p (20 consecutive entries, 128-bit loads; stride-21 rows)
                    {
                        P row[20];
                        load_vec<P, 20>(tcd + ((uint64_t) p0c + lane) * 20, row);
#pragma unroll
                        for (int t = 0; t < 20; t++) qs[lane * 21 + t] = row[t];
                        qs[lane * 21 + 20] = 0;                           // the 21st entry of a row serves "no code": piece 0
                    }
                    P qw[4];
                    load_vec<P, 4>(tw + p0c + pl, qw);
                    __syncwarp();
                    int nq = 0;                                           // queued items (warp-uniform)
                    auto flush = [&]() {
                        for (int q0 = 0; q0 < nq; q0 += 32) {
                            if (q0 + lane < nq) {
                                const uint32_t it = queue[q0 + lane], rr = it >> 5, pp = it & 31u;
                                const uint32_t rid = (uint32_t) ids[rr];
                                P fc[20], fq[20], e[20];
                                load_vec<P, 20>(s.vecs + ((uint64_t) (rid - nSeqs) * Lp + p0c + pp) * 20, fc);
                                load_vec<P, 20>(tv + (uint64_t) (p0c + pp) * 20, fq);
#pragma unroll
                                for (int i = 0; i < 20; i++) e[i] = s.eigenval[i];
                                // MODE 1: profileDist(node, out-profile) -- the candidate is the first profile; the first
                                // product commutes, so both argument orders give the same bits (NJ.tcc:914-916)
                                Pc[rr * WS + pp] = vec_mul3_sum<P, 20>(fq, fc, e, s.reduction);
                            }
                        }
                        nq = 0;
                        __syncwarp();
                    };
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int rr = u * 4 + sub;
                        const bool leafRow = (uint32_t) __shfl_sync(full, id, rr) < nSeqs;      // (dead rows: id -1 -> not a leaf, weights 0)
                        P wt[4], pc[4];
                        unsigned needMask[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t c = min((c4[u] >> (8 * i)) & 0xFFu, 20u);        // 20 = no code
                            // NJ.tcc:1174-1176: weight = p1->weights[i] * p2->weights[i] where both are > 0 (the product commutes).
                            // Weights are never negative, so the product is +0.0 exactly where the reference skips the position
                            // (and where it underflows the reference adds the same 0): no select needed; a leaf's weight is 1 or 0.
                            wt[i] = leafRow ? (c < 20u ? qw[i] : (P) 0) : pmul(w4[u][i], qw[i]);
                            pc[i] = qs[(pl + i) * 21 + c];                                  // (times wt == 0 where the position is off)
                            needMask[i] = __ballot_sync(full, c == 20u && wt[i] > 0);
                        }
                        if constexpr (sizeof(P) == 4) {
                            *reinterpret_cast<float4 *>(Wt + rr * WS + pl) = make_float4(wt[0], wt[1], wt[2], wt[3]);
                            *reinterpret_cast<float4 *>(Pc + rr * WS + pl) = make_float4(pc[0], pc[1], pc[2], pc[3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; i += 2) {
                                *reinterpret_cast<double2 *>(Wt + rr * WS + pl + i) = make_double2(wt[i], wt[i + 1]);
                                *reinterpret_cast<double2 *>(Pc + rr * WS + pl + i) = make_double2(pc[i], pc[i + 1]);
                            }
                        }
                        if (needMask[0] | needMask[1] | needMask[2] | needMask[3]) {
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                if (needMask[i] >> lane & 1u) queue[nq + __popc(needMask[i] & ((1u << lane) - 1u))] = (uint16_t) ((rr << 5) | (pl + i));
                                nq += __popc(needMask[i]);
                            }
                            if (nq > SWP_QCAP - 128) { __syncwarp(); flush(); }
                        }
                    }
                    __syncwarp();
                    if (nq > 0) flush();
                }
            } else if (r > 0 && id >= 0) {
                // ---- consumer: the tiles of round r-1, in chunk order; lane k adds row k in position order ----------------
                const int nT = min(NP, nChunks - (r - 1) * NP);
                for (int t = 0; t < nT; t++) {
                    const P *wr = tiles + ((size_t) ((r - 1) & 1) * NP + t) * (Cfg::tile / sizeof(P)) + lane * WS, *pr = wr + 32 * WS;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        P a[4], b[4];
                        load_vec<P, 4>(wr + j, a);
                        load_vec<P, 4>(pr + j, b);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const double wd = (double) a[i];
                            den = xadd(den, wd);                          // :1177 (skipped positions add +0.0: exact)
                            top = xadd(top, xmul(wd, (double) b[i]));    // :1183
                        }
                    }
                }
            }
            __syncthreads();                                              // round r's tiles complete; round r-1's consumed
        }
        if (!consumer || id < 0) continue;
        P dd, ww;
        finish_dist<P>(den, top, dd, ww);
        if (MODE == 1) {
            const P v = out_distance_finish<P>(s, id, nActive, totdiam, dd, ww);
            if (res != nullptr) res[k] = v; else s.outDist[id] = v;
            continue;
        }
        const int64_t a = MODE == 2 ? (int64_t) reqA[k] : query;
        P d, w;
        if (a < s.nSeqs && id < (int32_t) nSeqs) {                        // leaf x leaf: seqDist (NJ.tcc:1601-1624) -- the same terms,
            d = (P) xadd((double) dd, 0.0); w = den > 0 ? ww : (P) 0;     // weight nUse (0 without overlap), :1122
        } else { d = join_correct<P>(s, a, id, dd); w = ww; }
        if (MODE == 2) { dist[k] = d; weight[k] = w; continue; }
        const double outI = (double) s.outDist[query], outJ = (double) s.outDist[id];
        const P c = (P) xsub((double) d, xadd(outI, outJ) / (double) (nActive - 2));      // setCriterion, fresh out-distances (:1099-1107)
        dist[k] = d; weight[k] = w; crit[k] = c;
        keys[k] = order_key(c);
    }
}

static inline unsigned sweep_grid(int64_t nSlots) {
    const int64_t groups = (nSlots + 31) / 32;
    return (unsigned) std::max<int64_t>(1, std::min<int64_t>(groups, 148 * 12));
}

}  // namespace
