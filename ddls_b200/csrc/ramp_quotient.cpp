// ramp_quotient.cpp -- symmetry quotient of a lowered job (host-only; applied by ramp_register_template).
//
// Partitioning an op n ways (agents/partitioners/utils.py:42-110) makes n sub-ops with the same cost, the same parents and
// one server each, and n(n-1) flows per original edge with the same run time: in _run_lookahead (RCE:379-467) they become
// ready, tick and complete in the same ticks, with bit-identical remaining times.  This pass finds that structure in ANY
// lowered job by colour refinement (1-WL) and emits an equivalent, smaller job that the lookahead kernels simulate instead:
//
//   op class  = ops with equal cost and n_parents, on workers that carry the same classes in the same priority order (and
//               at the same position there), whose in- and out-deps fall into the same dep classes with the same
//               multiplicities.  The members of a class sit on distinct workers, so when the class wins its workers (RCE:44-67)
//               every member is its worker's winner: the trace's active-worker count adds the class size (RCE:709-715).
//   dep class = deps with equal run time and flow flag whose source / destination ops fall into the same classes.  Every
//               ready flow is ticked every tick whatever its priority (RCE:733-775), so a dep's state never depends on
//               its channel or priority; those only feed  t_comm = min over channels of remaining(winner)  (RCE:608-663).
//   channel group = channels that carry the same dep classes in the same priority order, where deps that leave the same
//               op class with the same run time ("twins": always ready in the same ticks with the same remaining time)
//               count as one.  When the groups' orders are restrictions of ONE order (priorities from one global ranking, no
//               cross-twin ties) a dep class is ONE entry that carries the set of groups its members lie on and its twins'
//               global rank; otherwise it is split into one entry per group, keyed by the twins' rank on that group.  On a
//               group the winner is the ready entry with the largest key among those whose set contains the group.
//   counters  = len(parent_deps_completed) of JOB:530 is kept per op CLASS, scaled by the class size: an entry adds its
//               member count, the class is readied when the count passes through n_parents x class size -- all entries of
//               a dep class complete in the same tick, so that is the tick in which every member's own count passes
//               through n_parents (the `==` of JOB:531 fires once per op).
//
// The refinement works on 64-bit hashes; the result is then VERIFIED exactly (every condition above, member by member).
// A job without symmetry comes back unchanged in size; the caller then registers the original.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "../../include/ramp_b200.h"

namespace {

inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// injective in h for fixed v and in v for fixed h (odd multiplier, bijective mix)
inline uint64_t combine(uint64_t h, uint64_t v) { return mix(h * 0xD1342543DE82EF95ull + mix(v)); }

inline uint64_t bits_of(double x) { x = x + 0.0; uint64_t b; memcpy(&b, &x, 8); return b; }

// dense ids in order of first appearance; returns the number of distinct values
int32_t relabel(const std::vector<uint64_t>& sig, std::vector<int32_t>& out) {
    std::unordered_map<uint64_t, int32_t> ids;
    ids.reserve(sig.size() * 2 + 16);
    out.resize(sig.size());
    for (size_t i = 0; i < sig.size(); ++i) {
        auto it = ids.find(sig[i]);
        if (it == ids.end()) it = ids.emplace(sig[i], (int32_t)ids.size()).first;
        out[i] = it->second;
    }
    return (int32_t)ids.size();
}

void rank_keys(const int64_t* prio, int32_t n, std::vector<uint32_t>& key) {   // as ramp_engine.cu make_rank_keys (RCE:56-66, 672-685)
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return prio[a] > prio[b]; });
    key.resize(n);
    for (int32_t r = 0; r < n; ++r) key[order[r]] = (uint32_t)(n - r);
}

template <class T> T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

void ramp_free_quotient(ramp_quotient_t* q) {
    if (!q) return;
    free(q->op_cost); free(q->op_key); free(q->op_worker); free(q->op_weight); free(q->op_threshold); free(q->row_ptr);
    free(q->dep_dst); free(q->dep_run_time); free(q->dep_key); free(q->dep_channel); free(q->dep_group_mask); free(q->dep_is_flow); free(q->dep_inc);
    free(q->op_class); free(q->dep_entry);
    memset(q, 0, sizeof(*q));
}

int ramp_quotient_template(const ramp_lowered_job_t* j, ramp_quotient_t* out) {
    if (!j || !out) return RAMP_ERR_BAD_ARG;
    memset(out, 0, sizeof(*out));
    const int32_t N = j->n_ops, E = j->n_deps, W = j->n_workers, C = j->n_channels;
    if (N < 1 || E < 0 || W < 1 || C < 0) return RAMP_ERR_BAD_ARG;
    std::vector<uint32_t> op_key, dep_key;
    rank_keys(j->op_prio, N, op_key);
    rank_keys(j->dep_prio, E, dep_key);
    std::vector<int32_t> src(E);
    for (int32_t i = 0; i < N; ++i) for (int32_t e = j->row_ptr[i]; e < j->row_ptr[i + 1]; ++e) src[e] = i;
    // in-CSR
    std::vector<int32_t> in_ptr(N + 1, 0), in_dep(E);
    for (int32_t e = 0; e < E; ++e) in_ptr[j->dep_dst[e] + 1]++;
    for (int32_t i = 0; i < N; ++i) in_ptr[i + 1] += in_ptr[i];
    { std::vector<int32_t> fill(in_ptr.begin(), in_ptr.end() - 1); for (int32_t e = 0; e < E; ++e) in_dep[fill[j->dep_dst[e]]++] = e; }
    // ops of every worker in descending key order; position of each op on its worker
    std::vector<int32_t> w_ptr(W + 1, 0), w_ops(N), w_pos(N);
    for (int32_t i = 0; i < N; ++i) { if (j->op_worker[i] >= W) return RAMP_ERR_BAD_ARG; w_ptr[j->op_worker[i] + 1]++; }
    for (int32_t w = 0; w < W; ++w) w_ptr[w + 1] += w_ptr[w];
    {
        std::vector<int32_t> order(N);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return op_key[a] > op_key[b]; });
        std::vector<int32_t> fill(w_ptr.begin(), w_ptr.end() - 1);
        for (int32_t i : order) { const int32_t w = j->op_worker[i]; w_pos[i] = fill[w] - w_ptr[w]; w_ops[fill[w]++] = i; }
    }

    // ---- colour refinement ----
    std::vector<int32_t> oc, dc, oc2, dc2;
    std::vector<uint64_t> sig;
    sig.resize(N);
    for (int32_t i = 0; i < N; ++i) sig[i] = combine(bits_of(j->op_cost[i]), (uint64_t)j->op_n_parents[i]);
    int32_t n_oc = relabel(sig, oc);
    sig.resize(E);
    for (int32_t e = 0; e < E; ++e) sig[e] = combine(bits_of(j->dep_run_time[e]), j->dep_is_flow[e] ? 1 : 0);
    int32_t n_dc = relabel(sig, dc);
    std::vector<uint64_t> wsig(W), tmp;
    for (int round = 0; round < 256; ++round) {
        for (int32_t w = 0; w < W; ++w) {
            uint64_t h = 0x1234567ull;
            for (int32_t k = w_ptr[w]; k < w_ptr[w + 1]; ++k) h = combine(h, (uint64_t)oc[w_ops[k]]);
            wsig[w] = h;
        }
        sig.resize(E);
        for (int32_t e = 0; e < E; ++e) sig[e] = combine(combine((uint64_t)dc[e], (uint64_t)oc[src[e]]), (uint64_t)oc[j->dep_dst[e]]);
        const int32_t n_dc2 = relabel(sig, dc2);
        sig.resize(N);
        for (int32_t i = 0; i < N; ++i) {
            uint64_t h = combine(combine((uint64_t)oc[i], wsig[j->op_worker[i]]), (uint64_t)w_pos[i]);
            tmp.clear();
            for (int32_t e = j->row_ptr[i]; e < j->row_ptr[i + 1]; ++e) tmp.push_back((uint64_t)dc2[e]);
            std::sort(tmp.begin(), tmp.end());
            for (uint64_t v : tmp) h = combine(h, v);
            h = combine(h, 0xABCDull);
            tmp.clear();
            for (int32_t k = in_ptr[i]; k < in_ptr[i + 1]; ++k) tmp.push_back((uint64_t)dc2[in_dep[k]]);
            std::sort(tmp.begin(), tmp.end());
            for (uint64_t v : tmp) h = combine(h, v);
            sig[i] = h;
        }
        const int32_t n_oc2 = relabel(sig, oc2);
        if (getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] round %d: op classes %d -> %d, dep classes %d -> %d\n", round, n_oc, n_oc2, n_dc, n_dc2);
        const bool stable = (n_oc2 == n_oc) && (n_dc2 == n_dc);
        oc.swap(oc2); dc.swap(dc2); n_oc = n_oc2; n_dc = n_dc2;
        if (stable) break;
    }
    // classes are numbered in order of their first member (relabel), so class 0 holds op 0 etc.

    // ---- exact verification of the op / dep classes ----
    std::vector<int32_t> rep_op(n_oc, -1), size_op(n_oc, 0), rep_dep(n_dc, -1);
    for (int32_t i = 0; i < N; ++i) { if (rep_op[oc[i]] < 0) rep_op[oc[i]] = i; size_op[oc[i]]++; }
    for (int32_t e = 0; e < E; ++e) if (rep_dep[dc[e]] < 0) rep_dep[dc[e]] = e;
    bool ok = true;
    for (int32_t e = 0; e < E && ok; ++e) {
        const int32_t r = rep_dep[dc[e]];
        ok = bits_of(j->dep_run_time[e]) == bits_of(j->dep_run_time[r]) && (!j->dep_is_flow[e]) == (!j->dep_is_flow[r])
             && oc[src[e]] == oc[src[r]] && oc[j->dep_dst[e]] == oc[j->dep_dst[r]];
        if (!ok && getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] dep %d vs rep %d\n", e, r);
    }
    {
        std::vector<int32_t> a, b;
        std::vector<int32_t> seen_worker(W, -1);       // class that last claimed the worker
        for (int32_t i = 0; i < N && ok; ++i) {
            const int32_t c = oc[i], r = rep_op[c];
            if (i == r) continue;
            ok = bits_of(j->op_cost[i]) == bits_of(j->op_cost[r]) && j->op_n_parents[i] == j->op_n_parents[r] && w_pos[i] == w_pos[r];
            if (!ok) { if (getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] op %d vs rep %d: cost/np/pos %d %d\n", i, r, w_pos[i], w_pos[r]); break; }
            const int32_t wi = j->op_worker[i], wr = j->op_worker[r];
            if (wi == wr) { if (getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] ops %d and %d of class %d share worker %d\n", i, r, c, wi); ok = false; break; }                                   // two members on one worker
            if (w_ptr[wi + 1] - w_ptr[wi] != w_ptr[wr + 1] - w_ptr[wr]) { ok = false; break; }
            for (int32_t k = 0; k < w_ptr[wi + 1] - w_ptr[wi] && ok; ++k) ok = oc[w_ops[w_ptr[wi] + k]] == oc[w_ops[w_ptr[wr] + k]];
            if (!ok) break;
            a.clear(); b.clear();
            for (int32_t e = j->row_ptr[i]; e < j->row_ptr[i + 1]; ++e) a.push_back(dc[e]);
            for (int32_t e = j->row_ptr[r]; e < j->row_ptr[r + 1]; ++e) b.push_back(dc[e]);
            std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
            if (a != b) { if (getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] op %d vs rep %d: out multiset\n", i, r); ok = false; break; }
            a.clear(); b.clear();
            for (int32_t k = in_ptr[i]; k < in_ptr[i + 1]; ++k) a.push_back(dc[in_dep[k]]);
            for (int32_t k = in_ptr[r]; k < in_ptr[r + 1]; ++k) b.push_back(dc[in_dep[k]]);
            std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
            if (a != b) { if (getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] op %d vs rep %d: in multiset\n", i, r); ok = false; break; }
        }
        (void)seen_worker;
    }
    if (!ok && getenv("RAMP_QUOTIENT_DEBUG")) fprintf(stderr, "[quotient] verification failed -> identity\n");
    if (!ok) {
        // (a 64-bit hash collision) fall back to the identity partition, which is trivially sound
        n_oc = N; n_dc = E;
        for (int32_t i = 0; i < N; ++i) { oc[i] = i; }
        for (int32_t e = 0; e < E; ++e) { dc[e] = e; }
        rep_op.resize(N); size_op.assign(N, 1); rep_dep.resize(E);
        std::iota(rep_op.begin(), rep_op.end(), 0); std::iota(rep_dep.begin(), rep_dep.end(), 0);
    }

    // ---- worker groups: workers with the same class sequence; the group's first worker lends the keys ----
    std::vector<int32_t> wg(W);
    int32_t n_wg = 0;
    {
        std::unordered_map<uint64_t, std::vector<int32_t>> by_hash;    // hash -> group ids, compared exactly
        std::vector<int32_t> first_worker;
        for (int32_t w = 0; w < W; ++w) {
            uint64_t h = 0x777ull;
            for (int32_t k = w_ptr[w]; k < w_ptr[w + 1]; ++k) h = combine(h, (uint64_t)oc[w_ops[k]]);
            int32_t g = -1;
            for (int32_t cand : by_hash[h]) {
                const int32_t f = first_worker[cand];
                if (w_ptr[f + 1] - w_ptr[f] != w_ptr[w + 1] - w_ptr[w]) continue;
                bool same = true;
                for (int32_t k = 0; k < w_ptr[w + 1] - w_ptr[w] && same; ++k) same = oc[w_ops[w_ptr[f] + k]] == oc[w_ops[w_ptr[w] + k]];
                if (same) { g = cand; break; }
            }
            if (g < 0) { g = n_wg++; first_worker.push_back(w); by_hash[h].push_back(g); }
            wg[w] = g;
        }
        // key of a class = key of its member on the group's first worker
        out->op_key = (uint32_t*)malloc(sizeof(uint32_t) * n_oc);
        out->op_worker = (uint32_t*)malloc(sizeof(uint32_t) * n_oc);
        std::vector<char> have(n_oc, 0);
        for (int32_t g = 0; g < n_wg; ++g) {
            const int32_t f = first_worker[g];
            for (int32_t k = w_ptr[f]; k < w_ptr[f + 1]; ++k) {
                const int32_t i = w_ops[k];
                out->op_key[oc[i]] = op_key[i];
                out->op_worker[oc[i]] = (uint32_t)g;
                have[oc[i]] = 1;
            }
        }
        for (int32_t c = 0; c < n_oc; ++c) if (!have[c]) { ramp_free_quotient(out); return RAMP_ERR_BAD_ARG; }   // cannot happen
    }

    // ---- channel groups over twin-reduced sequences ----
    // twin id of a flow = (class of its source op, run time): twins are ready in the same ticks with the same remaining time
    std::vector<int32_t> twin(E, -1);
    {
        sig.resize(E);
        for (int32_t e = 0; e < E; ++e) sig[e] = combine((uint64_t)oc[src[e]], combine(bits_of(j->dep_run_time[e]), j->dep_is_flow[e] ? 1 : 0));
        std::vector<int32_t> t;
        relabel(sig, t);
        // exactness of the twin relation does not rest on the hash: verified against the first member below
        std::vector<int32_t> first(E, -1);
        for (int32_t e = 0; e < E; ++e) {
            int32_t& f = first[t[e]];
            if (f < 0) f = e;
            if (oc[src[e]] != oc[src[f]] || bits_of(j->dep_run_time[e]) != bits_of(j->dep_run_time[f]) || (!j->dep_is_flow[e]) != (!j->dep_is_flow[f])) {
                t[e] = -2 - e;      // collision: make it its own twin class (negative ids are unique)
            }
        }
        twin = t;
    }
    std::vector<int32_t> cg(std::max(C, 1), -1);
    int32_t n_cg = 0;
    std::vector<int32_t> c_first;                                  // first channel of every group
    std::vector<std::vector<std::pair<int32_t, uint32_t>>> c_seq;  // per group: (twin, max key on the first channel), descending key
    {
        std::vector<std::vector<int32_t>> on_chan(C);
        for (int32_t e = 0; e < E; ++e)
            if (j->dep_channel[e] != RAMP_NO_CHANNEL) { if (j->dep_channel[e] >= C) { ramp_free_quotient(out); return RAMP_ERR_BAD_ARG; } on_chan[j->dep_channel[e]].push_back(e); }
        std::unordered_map<uint64_t, std::vector<int32_t>> by_hash;
        std::vector<std::pair<int32_t, uint32_t>> seq;
        for (int32_t c = 0; c < C; ++c) {
            std::unordered_map<int32_t, uint32_t> best;
            for (int32_t e : on_chan[c]) { uint32_t& b = best[twin[e]]; b = std::max(b, dep_key[e]); }
            seq.assign(best.begin(), best.end());
            std::sort(seq.begin(), seq.end(), [](const std::pair<int32_t, uint32_t>& a, const std::pair<int32_t, uint32_t>& b) { return a.second > b.second; });
            uint64_t h = 0x999ull;
            for (auto& p : seq) h = combine(h, (uint64_t)(uint32_t)p.first);
            int32_t g = -1;
            for (int32_t cand : by_hash[h]) {
                const auto& s2 = c_seq[cand];
                if (s2.size() != seq.size()) continue;
                bool same = true;
                for (size_t k = 0; k < seq.size() && same; ++k) same = s2[k].first == seq[k].first;
                if (same) { g = cand; break; }
            }
            if (g < 0) { g = n_cg++; c_first.push_back(c); c_seq.push_back(seq); by_hash[h].push_back(g); }
            cg[c] = g;
        }
    }
    // key of (twin, group): looked up through a per-group map
    std::vector<std::unordered_map<int32_t, uint32_t>> c_key(n_cg);
    for (int32_t g = 0; g < n_cg; ++g) for (auto& p : c_seq[g]) c_key[g][p.first] = p.second;

    // ---- one key per twin for ALL groups?  The global rank of a twin = the largest original key among its deps.  If every
    // group's sequence is strictly descending in it, the per-channel orders are restrictions of ONE order (they are whenever the
    // priorities come from one global ranking by run time, srpt_dep_scheduler.py:60-81, with no cross-twin ties), and a dep class
    // needs a single entry carrying the SET of groups its members lie on: on group g the winner is the live entry with the
    // largest key among those whose set contains g.  Otherwise (hash-ordered ties) a class is split into one entry per group. ----
    std::unordered_map<int32_t, uint32_t> gk;              // twin -> global key
    for (int32_t e = 0; e < E; ++e) if (j->dep_channel[e] != RAMP_NO_CHANNEL) { uint32_t& b = gk[twin[e]]; b = std::max(b, dep_key[e]); }
    bool merged = n_cg <= 64;
    for (int32_t g = 0; g < n_cg && merged; ++g)
        for (size_t k = 1; k < c_seq[g].size() && merged; ++k) merged = gk[c_seq[g][k - 1].first] > gk[c_seq[g][k].first];
    {
        // two twins with the same global key would compare equal: impossible (keys are unique ranks), checked for safety
        std::vector<uint32_t> all;
        for (auto& p : gk) all.push_back(p.second);
        std::sort(all.begin(), all.end());
        for (size_t k = 1; k < all.size() && merged; ++k) merged = all[k] != all[k - 1];
    }
    const bool masks_valid = n_cg <= 64;

    // ---- entries: merged: one per dep class; split: one per (dep class, channel group or none) ----
    std::vector<int32_t> entry(E);
    std::vector<int32_t> ent_rep;         // representative dep of every entry
    std::vector<uint32_t> ent_inc;
    std::vector<uint64_t> ent_mask;
    {
        std::unordered_map<uint64_t, int32_t> ids;
        ids.reserve((size_t)E * 2 + 16);
        for (int32_t e = 0; e < E; ++e) {
            const int32_t g = (j->dep_channel[e] == RAMP_NO_CHANNEL) ? -1 : cg[j->dep_channel[e]];
            const uint64_t k = ((uint64_t)(uint32_t)dc[e] << 32) | (merged ? (g < 0 ? 0ull : 1ull) : (uint64_t)(uint32_t)(g + 1));
            auto it = ids.find(k);
            if (it == ids.end()) { it = ids.emplace(k, (int32_t)ent_rep.size()).first; ent_rep.push_back(e); ent_inc.push_back(0); ent_mask.push_back(0); }
            entry[e] = it->second;
            ent_inc[it->second]++;
            if (g >= 0 && masks_valid) ent_mask[it->second] |= (1ull << g);
        }
    }
    const int32_t n_ent = (int32_t)ent_rep.size();
    // CSR by source class, entries in order of their first member
    std::vector<int32_t> order(n_ent);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return oc[src[ent_rep[a]]] < oc[src[ent_rep[b]]]; });
    std::vector<int32_t> new_id(n_ent);
    for (int32_t k = 0; k < n_ent; ++k) new_id[order[k]] = k;

    out->n_ops = n_oc; out->n_deps = n_ent; out->n_workers = n_wg; out->n_channels = n_cg;
    out->op_cost = (double*)malloc(sizeof(double) * n_oc);
    out->op_weight = (uint32_t*)malloc(sizeof(uint32_t) * n_oc);
    out->op_threshold = (uint32_t*)malloc(sizeof(uint32_t) * n_oc);
    out->row_ptr = (int32_t*)calloc((size_t)n_oc + 1, sizeof(int32_t));
    out->dep_dst = (int32_t*)malloc(sizeof(int32_t) * std::max(n_ent, 1));
    out->dep_run_time = (double*)malloc(sizeof(double) * std::max(n_ent, 1));
    out->dep_key = (uint32_t*)malloc(sizeof(uint32_t) * std::max(n_ent, 1));
    out->dep_channel = (uint32_t*)malloc(sizeof(uint32_t) * std::max(n_ent, 1));
    out->dep_group_mask = (uint64_t*)malloc(sizeof(uint64_t) * std::max(n_ent, 1));
    out->merged = merged ? 1 : 0; out->masks_valid = masks_valid ? 1 : 0;
    out->dep_is_flow = (uint8_t*)malloc(std::max(n_ent, 1));
    out->dep_inc = (uint32_t*)malloc(sizeof(uint32_t) * std::max(n_ent, 1));
    out->op_class = (int32_t*)malloc(sizeof(int32_t) * N);
    out->dep_entry = (int32_t*)malloc(sizeof(int32_t) * std::max(E, 1));
    for (int32_t c = 0; c < n_oc; ++c) {
        out->op_cost[c] = j->op_cost[rep_op[c]] + 0.0;
        out->op_weight[c] = (uint32_t)size_op[c];
        out->op_threshold[c] = (uint32_t)j->op_n_parents[rep_op[c]] * (uint32_t)size_op[c];
    }
    for (int32_t k = 0; k < n_ent; ++k) {
        const int32_t id = order[k], e = ent_rep[id];
        out->row_ptr[oc[src[e]] + 1]++;
        out->dep_dst[k] = oc[j->dep_dst[e]];
        out->dep_run_time[k] = j->dep_run_time[e] + 0.0;
        out->dep_is_flow[k] = j->dep_is_flow[e] ? 1 : 0;
        out->dep_inc[k] = ent_inc[id];
        out->dep_group_mask[k] = ent_mask[id];
        if (j->dep_channel[e] == RAMP_NO_CHANNEL) {
            out->dep_channel[k] = 0xFFFFFFFFu;
            out->dep_key[k] = dep_key[e];
        } else if (merged) {
            out->dep_channel[k] = 0xFFFFFFFFu;                 // lies on the groups of dep_group_mask
            out->dep_key[k] = gk[twin[e]];
        } else {
            const int32_t g = cg[j->dep_channel[e]];
            out->dep_channel[k] = (uint32_t)g;
            out->dep_key[k] = c_key[g][twin[e]];
        }
    }
    for (int32_t c = 0; c < n_oc; ++c) out->row_ptr[c + 1] += out->row_ptr[c];
    memcpy(out->op_class, oc.data(), sizeof(int32_t) * N);
    for (int32_t e = 0; e < E; ++e) out->dep_entry[e] = new_id[entry[e]];
    return RAMP_OK;
}

}  // extern "C"
