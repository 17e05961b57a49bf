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
// Work decomposition (one warp = 32 candidates, "tile-transposed" like group_profile_dist, but 4 positions per lane):
//   per chunk of 32 positions   the query's [32][20] table slice is staged in shared memory (coalesced);
//     phase 1  8 units of 4 rows x 32 positions: lane l handles row 4u + l/8, positions 4(l%8)..+3 -- one 32-bit code load,
//              one 128-bit weight load, 4 shared-memory gathers -> w1*w2 and the piece go into a [32][32] tile;
//     queue    candidate positions with a vector: lane-per-item 20-wide products, results patched into the tile;
//     phase 2  lane r adds row r left to right into ITS candidate's denom / top (the reference's order: one dependent
//              DADD per position and chain, NJ.tcc:1172-1183; top's term w*piece is formed here in double).
// Bytes per candidate position: 1 (leaf) or 5 (internal node, fp32) -- the kernel is bound by L2/HBM bandwidth and the
// phase-2 chains, not by instruction issue as the generic 1-position-per-lane kernel was.
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

constexpr int SWP_QCAP = 256;                    // queued vector positions per warp (flushed above 128; a unit adds <= 128)
// R = candidates (tile rows) per warp: 32 when the sweep has enough candidates to fill the machine with 32-row warps, 16 or 8
// for the shorter sweeps (20 000 candidates are only 625 32-row warps: 4 per SM, one per scheduler -- latency bound).
// MODE 0/1: every warp of the CTA works on the same query, chunk by chunk in lock step, and the query's table slice is staged
// ONCE per CTA (double-buffered, one __syncthreads per chunk); MODE 2 (a query per list): per warp.
template<typename P, int MODE, int R> struct SweepCfg {
    static constexpr int NW = R == 32 ? 4 : 8;                 // warps per CTA
    static constexpr int WS = sizeof(P) == 4 ? 36 : 34;        // tile row stride: 16-byte aligned rows, conflict-free 128-bit row reads
    static constexpr size_t tile = 2 * (size_t) R * WS * sizeof(P);   // w1*w2 and piece
    static constexpr size_t qsOne = 32 * 21 * sizeof(P);       // one table slice, row stride 21
    static constexpr size_t perWarp = tile + SWP_QCAP * 2 + R * 4 + (MODE == 2 ? qsOne : 0);
    static constexpr size_t shared = MODE == 2 ? 0 : 2 * qsOne;
    static constexpr size_t bytes = shared + NW * perWarp;
};

template<typename P, int MODE, int R>
__global__ void __launch_bounds__(SweepCfg<P, MODE, R>::NW * 32, sizeof(P) == 4 ? (R == 32 ? 4 : 2) : 1)
k_sweep20(Store<P> s, QTab<P> qt, const int32_t *__restrict__ list, int stride, int offset,
          const int32_t *__restrict__ reqA, const int32_t *__restrict__ reqB, int cap,
          int64_t nSlots, int64_t query, int64_t nActive, double totdiam,
          P *__restrict__ dist, P *__restrict__ weight, P *__restrict__ crit, uint64_t *__restrict__ keys, P *__restrict__ res) {
    typedef SweepCfg<P, MODE, R> Cfg;
    static_assert(MODE != 2 || R == 32, "a query per list: 32-row warps");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int WS = Cfg::WS, NW = Cfg::NW, U = R / 4;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *sm = smemRaw + Cfg::shared + (size_t) wid * Cfg::perWarp;
    P *Wt = reinterpret_cast<P *>(sm);                                   // [R][WS]
    P *Pc = Wt + R * WS;                                                 // [R][WS]
    uint16_t *queue = reinterpret_cast<uint16_t *>(sm + Cfg::tile);
    int32_t *ids = reinterpret_cast<int32_t *>(sm + Cfg::tile + SWP_QCAP * 2);
    P *qsWarp = reinterpret_cast<P *>(sm + Cfg::tile + SWP_QCAP * 2 + R * 4);     // MODE 2 only
    P *qsCta = reinterpret_cast<P *>(smemRaw);                           // MODE 0/1: [2][32][21]
    const uint32_t Lp = (uint32_t) s.Lp, nSeqs = (uint32_t) s.nSeqs;
    const int nChunks = (int) (Lp / 32);
    const int64_t nGroups = (nSlots + R - 1) / R, nSuper = (nGroups + NW - 1) / NW;
    const int sub = lane >> 3, pl = (lane & 7) * 4;                      // my row within a unit, my first position within a chunk

    for (int64_t sg = blockIdx.x; sg < nSuper; sg += gridDim.x) {
        const int64_t g = sg * NW + wid;
        const int64_t k = g * R + lane;
        int32_t id = -1;
        if (lane < R && k < nSlots) {
            if (MODE == 2) { if (reqA[k] >= 0) id = reqB[k]; }
            else id = list[k * stride + offset];
        }
        const bool alive = __ballot_sync(full, id >= 0) != 0;            // (MODE 0/1: a dead warp still takes part in the staging)
        if (MODE == 2 && !alive) continue;
        const size_t tab = MODE == 2 ? (size_t) ((g * R) / cap) : 0;      // cap is a multiple of 32: one query per group
        const P *tcd = qt.cd + tab * qt.stride, *tv = qt.v + tab * qt.stride, *tw = qt.w + tab * (qt.stride / 20);
        if (lane < R) ids[lane] = id;
        __syncwarp();
        double den = 0, top = 0;
        int nq = 0;                                                       // queued items (warp-uniform)
        // flush: the queued (row, position) items of this chunk -- candidate positions that carry a vector
        auto flush = [&](uint32_t p0c) {
            for (int q0 = 0; q0 < nq; q0 += 32) {
                if (q0 + lane < nq) {
                    const uint32_t it = queue[q0 + lane], r = it >> 5, pp = it & 31u;
                    const uint32_t rid = (uint32_t) ids[r];
                    P fc[20], fq[20], e[20];
                    load_vec<P, 20>(s.vecs + ((uint64_t) (rid - nSeqs) * Lp + p0c + pp) * 20, fc);
                    load_vec<P, 20>(tv + (uint64_t) (p0c + pp) * 20, fq);
#pragma unroll
                    for (int i = 0; i < 20; i++) e[i] = s.eigenval[i];
                    // MODE 1: profileDist(node, out-profile) -- the candidate is the first profile; the first product
                    // commutes, so both argument orders give the same bits (NJ.tcc:914-916)
                    Pc[r * WS + pp] = vec_mul3_sum<P, 20>(fq, fc, e, s.reduction);
                }
            }
            nq = 0;
            __syncwarp();
        };
        // the query's [32][20] table slice of chunk ch -> shared memory (coalesced reads, row stride 21 against bank conflicts)
        auto stage = [&](int ch) {
            const P *src = tcd + (uint64_t) ch * 32 * 20;
            if constexpr (MODE == 2) {
#pragma unroll
                for (int t = 0; t < 20; t++) { const int idx = t * 32 + lane; qsWarp[(idx / 20) * 21 + (idx % 20)] = src[idx]; }
            } else {
                P *dst = qsCta + (size_t) (ch & 1) * 32 * 21;
                for (int idx = threadIdx.x; idx < 640; idx += NW * 32) dst[(idx / 20) * 21 + (idx % 20)] = src[idx];
            }
        };
        if constexpr (MODE != 2) { stage(0); __syncthreads(); }
        for (int ch = 0; ch < nChunks; ch++) {
            const uint32_t p0c = (uint32_t) ch * 32;
            const P *qs = qsWarp;
            if constexpr (MODE == 2) stage(ch);
            else { if (ch + 1 < nChunks) stage(ch + 1); qs = qsCta + (size_t) (ch & 1) * 32 * 21; }
            if (alive) {
                P qw[4];
                load_vec<P, 4>(tw + p0c + pl, qw);
                // phase 1 loads: codes + weights of my 4 positions in each of the U units
                uint32_t c4[U];
                P w4[U][4];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int32_t rid = __shfl_sync(full, id, u * 4 + sub);
                    c4[u] = 0x7F7F7F7Fu;
#pragma unroll
                    for (int i = 0; i < 4; i++) w4[u][i] = 0;
                    if (rid >= 0) {
                        c4[u] = *reinterpret_cast<const uint32_t *>(s.codes + (uint64_t) (uint32_t) rid * Lp + p0c + pl);
                        if ((uint32_t) rid >= nSeqs) load_vec<P, 4>(s.weights + (uint64_t) ((uint32_t) rid - nSeqs) * Lp + p0c + pl, w4[u]);
                        else {
#pragma unroll
                            for (int i = 0; i < 4; i++) w4[u][i] = ((c4[u] >> (8 * i)) & 0xFFu) != VFT_DEV_NOCODE ? (P) 1 : (P) 0;
                        }
                    }
                }
                if constexpr (MODE == 2) __syncwarp();                    // qs staged
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int r = u * 4 + sub;
                    P wt[4], pc[4];
                    unsigned needMask[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t c = (c4[u] >> (8 * i)) & 0xFFu;
                        const bool on = w4[u][i] > 0 && qw[i] > 0;
                        // NJ.tcc:1176: weight = p1->weights[i] * p2->weights[i] (the product commutes)
                        wt[i] = on ? pmul(w4[u][i], qw[i]) : (P) 0;
                        pc[i] = (on && c < 20u) ? qs[(pl + i) * 21 + c] : (P) 0;
                        needMask[i] = __ballot_sync(full, on && c == VFT_DEV_NOCODE);
                    }
                    if constexpr (sizeof(P) == 4) {
                        *reinterpret_cast<float4 *>(Wt + r * WS + pl) = make_float4(wt[0], wt[1], wt[2], wt[3]);
                        *reinterpret_cast<float4 *>(Pc + r * WS + pl) = make_float4(pc[0], pc[1], pc[2], pc[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i += 2) {
                            *reinterpret_cast<double2 *>(Wt + r * WS + pl + i) = make_double2(wt[i], wt[i + 1]);
                            *reinterpret_cast<double2 *>(Pc + r * WS + pl + i) = make_double2(pc[i], pc[i + 1]);
                        }
                    }
                    if (needMask[0] | needMask[1] | needMask[2] | needMask[3]) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if (needMask[i] >> lane & 1u) queue[nq + __popc(needMask[i] & ((1u << lane) - 1u))] = (uint16_t) ((r << 5) | (pl + i));
                            nq += __popc(needMask[i]);
                        }
                        if (nq > SWP_QCAP - 128) { __syncwarp(); flush(p0c); }
                    }
                }
                __syncwarp();
                if (nq > 0) flush(p0c);
                // phase 2: lane r adds row r in position order
                if (id >= 0) {
                    const P *wr = Wt + lane * WS, *pr = Pc + lane * WS;
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
                __syncwarp();
            }
            if constexpr (MODE != 2) __syncthreads();                     // slice ch+1 staged by everybody, slice ch free again
        }
        if (id < 0) continue;
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

// rows per warp of a sweep over n candidates: the largest of 32 / 16 / 8 that still gives ~16 warps per SM
static inline int sweep_rows(int64_t n) { return n >= 32 * 2368 ? 32 : (n >= 16 * 2368 ? 16 : 8); }
template<int R> static inline unsigned sweep_grid(int64_t nSlots) {
    constexpr int NW = R == 32 ? 4 : 8;
    const int64_t groups = (nSlots + R - 1) / R, super = (groups + NW - 1) / NW;
    return (unsigned) std::max<int64_t>(1, std::min<int64_t>(super, 148 * 8));
}

}  // namespace
