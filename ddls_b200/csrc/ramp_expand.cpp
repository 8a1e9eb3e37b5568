// ramp_expand.cpp -- native "template expansion" (SURVEY.md 8f-1): forward graph + partition degree + server block ->
// the lowered job _run_lookahead consumes, without the reference's Python object churn.
//
// Restates, for one job on a block of servers (sub-op k of every split op on the k-th server of the block):
//   mirrored forward/backward graph                                   ddls/utils.py:342-415
//   per-op split count                                                RJPE:332-343
//   model_split_node (sub-ops, fanned-out edges, sync edges, sizes)   agents/partitioners/utils.py:42-110
//   update_dep_run_times (collectives / one-to-one)                   actions/utils.py:13-393
//   SRPT op / dep priorities                                          srpt_op_scheduler.py:16-88, srpt_dep_scheduler.py:14-83
//   first-fit dep placement on the one-hop channel                    first_fit_dep_placer.py:23-160
// The Python twin is ddls_b200/template_builder.py (pinned against the reference's own lowered jobs in
// tests/test_lowering_roundtrip.py); tests/test_expand_native.py requires this file to equal it bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/ramp_b200.h"

int ramp_internal_set_error(int code, const char* msg);   // ramp_engine.cu: what ramp_last_error() returns

namespace {

struct Node {
    std::string id;
    double cost = 0.0, mem = 0.0;
    bool alive = true;
    std::vector<int> succ, pred;     // adjacency in insertion order (networkx semantics)
};

struct Graph {
    std::vector<Node> nodes;                         // creation order == dict order of the alive ones
    std::unordered_map<uint64_t, double> size;      // (u << 32 | v) -> edge size
    static uint64_t key(int u, int v) { return ((uint64_t)(uint32_t)u << 32) | (uint32_t)v; }
    int add_node(const std::string& id, double cost, double mem) {
        nodes.push_back(Node{id, cost, mem, true, {}, {}});
        return (int)nodes.size() - 1;
    }
    void add_edge(int u, int v, double sz) {
        auto it = size.find(key(u, v));
        if (it == size.end()) { nodes[u].succ.push_back(v); nodes[v].pred.push_back(u); size[key(u, v)] = sz; }
        else it->second = sz;
    }
    double get_size(int u, int v, double dflt) const { auto it = size.find(key(u, v)); return it == size.end() ? dflt : it->second; }
};

void erase_first(std::vector<int>& v, int x) {
    auto it = std::find(v.begin(), v.end(), x);
    if (it != v.end()) v.erase(it);
}

// calc_ramp_all_reduce_collective_communication_run_time, actions/utils.py:40-88 (same operations in the same order)
double all_reduce_time(double message, int node_ids, int racks, int cgs, int x, double data_rate, double latency, double io) {
    const double mem_frq = 2e12, peak = 130e12, bytes_per_comp = 2.0;
    auto trx = [&](double d) -> double {                         // effective_trx_per_comm(cg=x, d, J=1)
        if (d == 1.0) return 0.0;
        const double a = (double)x, b = std::floor((double)x / (d - 1.0));
        return 1.0 + (std::min(a, b) - 1.0);
    };
    auto add_time = [&](double data_sz, double devices) -> double {
        const double n_op = std::ceil(std::log2(devices));
        const double n_bytes = (devices + 1.0) * bytes_per_comp;
        const double ai = n_op / n_bytes;
        const double total_ops = n_op * (data_sz / devices) / bytes_per_comp;
        return total_ops / std::min(mem_frq * ai, peak);
    };
    const double data_per_tx = data_rate / (double)x;
    const double sub[4] = {(double)cgs, (double)std::min(cgs, node_ids), (double)racks, std::ceil((double)node_ids / (double)x)};
    double bw[4], msg[4];
    for (int i = 0; i < 4; ++i) bw[i] = trx(sub[i]) * data_per_tx;
    msg[0] = std::ceil(message / sub[0]);
    for (int i = 1; i < 4; ++i) msg[i] = std::ceil(msg[i - 1] / sub[i]);
    double comm = 0.0, comp = 0.0;
    for (int i = 0; i < 4; ++i) {
        if (sub[i] > 1.0) {
            comp += add_time(msg[i] * sub[i], sub[i]);
            comm += latency + 2.0 * io + msg[i] / bw[i];
        }
    }
    return 2.0 * comm + comp;
}

template <class T> T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

void ramp_free_expanded_aux(ramp_expanded_aux_t* a) {
    if (!a) return;
    free(a->dep_size); free(a->op_mem); free(a->node_order);
    memset(a, 0, sizeof(*a));
}

void ramp_free_expanded_job(ramp_lowered_job_t* j) {
    if (!j) return;
    free((void*)j->op_cost); free((void*)j->op_prio); free((void*)j->op_worker); free((void*)j->op_n_parents);
    free((void*)j->row_ptr); free((void*)j->dep_dst); free((void*)j->dep_run_time); free((void*)j->dep_prio);
    free((void*)j->dep_channel); free((void*)j->dep_is_flow);
    memset(j, 0, sizeof(*j));
}

int ramp_expand_template(const ramp_forward_graph_t* g, int32_t degree, double quantum, const ramp_block_t* blk,
                         int32_t run_time_mode, int32_t num_training_steps, ramp_lowered_job_t* out, ramp_expanded_aux_t* aux) {
    if (!g || !blk || !out || g->n_fwd < 1 || degree < 1 || (degree != 1 && degree % 2 != 0) || blk->n_servers < 1)
        return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_expand_template: null argument, empty graph or block, or a degree that is neither 1 nor even (op_partition.py:26-27)");
    const int n = g->n_fwd;
    Graph G;
    {   // every forward edge fans out to at most degree^2 edges per direction, every split backward op adds degree^2 sync edges
        const size_t est = (size_t)(2 * g->n_edges + 1 + n) * (size_t)degree * (size_t)degree + 16;
        G.size.reserve(est * 2);
        G.nodes.reserve((size_t)2 * n * (degree + 1) + 4);
    }
    std::vector<int> fwd_node(n + 1), bwd_node(n + 1);
    std::vector<double> mem0(n + 1);
    for (int i = 1; i <= n; ++i) {                               // utils.py:432 memory_cost = activation + parameter
        mem0[i] = g->act_size[i - 1] + g->par_size[i - 1];
        fwd_node[i] = G.add_node(std::to_string(i), g->fwd_cost[i - 1], mem0[i]);
        bwd_node[i] = G.add_node(std::to_string(2 * n - (i - 1)), g->bwd_cost[i - 1], mem0[i]);
    }
    for (int e = 0; e < g->n_edges; ++e) {
        const int u = g->edge_src[e], v = g->edge_dst[e];
        if (u < 1 || u > n || v < 1 || v > n) return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_expand_template: an edge names an op outside 1..n_fwd");
        G.add_edge(fwd_node[u], fwd_node[v], G.nodes[fwd_node[u]].mem);
    }
    for (int e = 0; e < g->n_edges; ++e) {                       // mirrored backward edge 2n-(v-1) -> 2n-(u-1)
        const int u = g->edge_src[e], v = g->edge_dst[e];
        G.add_edge(bwd_node[v], bwd_node[u], G.nodes[bwd_node[v]].mem);
    }
    G.add_edge(fwd_node[n], bwd_node[n], G.nodes[fwd_node[n]].mem);    // join edge n -> n+1

    // ---- model_split_node ----
    std::vector<int> splits(n + 1, 1);
    for (int i = 1; i <= n; ++i) {
        const double c = g->fwd_cost[i - 1];
        const double k = std::max(1.0, std::min(std::ceil(std::ceil(c / quantum) / 2.0) * 2.0, (double)degree));   // RJPE:336
        splits[i] = (int)k;
        if (splits[i] > blk->n_servers)
            return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_expand_template: an op splits into more sub-ops than the block has servers");
        // the block holds one server per sub-op of the most-split op (which may be < degree)
    }
    std::unordered_map<uint64_t, double> in_feat, out_feat;
    in_feat.reserve(G.size.bucket_count()); out_feat.reserve(G.size.bucket_count());
    std::vector<uint64_t> in_order, out_order;                   // insertion order is irrelevant for the override below
    std::vector<std::vector<int>> fsubs(n + 1), bsubs(n + 1);
    for (int i = 1; i <= n; ++i) {
        const int k = splits[i];
        if (k <= 1) continue;
        for (int which = 0; which < 2; ++which) {
            const int node = which == 0 ? fwd_node[i] : bwd_node[i];
            const std::vector<int> ins = G.nodes[node].pred, outs = G.nodes[node].succ;
            const double sub_cost = G.nodes[node].cost / k, sub_mem = G.nodes[node].mem / k;
            for (int p : ins) { erase_first(G.nodes[p].succ, node); G.size.erase(Graph::key(p, node)); }
            for (int s : outs) { erase_first(G.nodes[s].pred, node); G.size.erase(Graph::key(node, s)); }
            G.nodes[node].alive = false;
            std::vector<int> subs;
            for (int j = 0; j < k; ++j) subs.push_back(G.add_node(G.nodes[node].id + (char)('a' + j), sub_cost, sub_mem));
            for (int sid : subs) {
                for (int p : ins) { G.add_edge(p, sid, G.get_size(p, sid, 0.0)); in_feat[Graph::key(p, sid)] = G.nodes[p].mem / k; }
                for (int s : outs) { G.add_edge(sid, s, G.get_size(sid, s, 0.0)); out_feat[Graph::key(sid, s)] = G.nodes[s].mem / k; }
            }
            if (which == 1)                                      // weight-sync collective between the backward sub-ops
                for (int a : subs) for (int b : subs) if (a != b) { G.add_edge(a, b, 0.0); in_feat[Graph::key(a, b)] = sub_mem; }
            (which == 0 ? fsubs[i] : bsubs[i]) = subs;
        }
    }
    for (auto& kv : in_feat) { auto it = G.size.find(kv.first); if (it != G.size.end()) it->second = kv.second; }
    for (auto& kv : out_feat) { auto it = G.size.find(kv.first); if (it != G.size.end()) it->second = kv.second; }

    // ---- op index = rank of the id string (RCE:56) ----
    std::vector<int> alive;
    for (int h = 0; h < (int)G.nodes.size(); ++h) if (G.nodes[h].alive) alive.push_back(h);      // dict order
    std::vector<int> by_id = alive;
    std::sort(by_id.begin(), by_id.end(), [&](int a, int b) { return G.nodes[a].id < G.nodes[b].id; });
    const int N = (int)by_id.size();
    std::vector<int> idx(G.nodes.size(), -1);
    for (int i = 0; i < N; ++i) idx[by_id[i]] = i;
    std::vector<double> op_cost(N);
    std::vector<int> local_worker(N);
    for (int i = 0; i < N; ++i) {
        const Node& nd = G.nodes[by_id[i]];
        op_cost[i] = nd.cost;
        const char last = nd.id.back();
        local_worker[i] = (last >= 'a' && last <= 'z') ? (last - 'a') : 0;
    }
    std::vector<int> used = local_worker;
    std::sort(used.begin(), used.end());
    used.erase(std::unique(used.begin(), used.end()), used.end());
    const int W = (int)used.size();
    std::vector<int> remap(*std::max_element(used.begin(), used.end()) + 1, -1);
    for (int i = 0; i < W; ++i) remap[used[i]] = i;
    std::vector<uint16_t> op_worker(N);
    for (int i = 0; i < N; ++i) op_worker[i] = (uint16_t)remap[local_worker[i]];

    // ---- deps in sorted (u, v) order == CSR by source ----
    struct Dep { int u, v; double size; };
    std::vector<Dep> deps;
    for (int i = 0; i < N; ++i) {
        const Node& nd = G.nodes[by_id[i]];
        std::vector<int> sv = nd.succ;
        std::sort(sv.begin(), sv.end(), [&](int a, int b) { return G.nodes[a].id < G.nodes[b].id; });
        for (int v : sv) deps.push_back(Dep{by_id[i], v, G.size[Graph::key(by_id[i], v)]});
    }
    const int E = (int)deps.size();
    std::vector<int32_t> row_ptr(N + 1, 0), dep_dst(E);
    std::vector<double> sizes(E), run_time(E), sched_cost(E);
    std::vector<uint8_t> is_flow(E);
    std::unordered_map<uint64_t, int> dep_index;
    dep_index.reserve((size_t)E * 2);
    for (int e = 0; e < E; ++e) {
        row_ptr[idx[deps[e].u] + 1]++;
        dep_dst[e] = idx[deps[e].v];
        sizes[e] = deps[e].size;
        dep_index[Graph::key(deps[e].u, deps[e].v)] = e;
    }
    for (int i = 0; i < N; ++i) row_ptr[i + 1] += row_ptr[i];
    const double one_to_one_base = blk->latency + 2 * blk->io_latency;
    for (int e = 0; e < E; ++e) {
        const int sw = op_worker[idx[deps[e].u]], dw = op_worker[dep_dst[e]];
        is_flow[e] = (sw != dw && sizes[e] != 0.0) ? 1 : 0;       // RCE:531-536 (one worker per server RCE:180)
        run_time[e] = is_flow[e] ? one_to_one_base + sizes[e] / blk->channel_bandwidth : 0.0;
    }

    // ---- update_dep_run_times with the reference's collective formulas ----
    const bool reference_times = run_time_mode == RAMP_RUN_TIMES_REFERENCE;
    if (reference_times) {
        auto coord = [&](int h) { const int w = op_worker[idx[h]]; return std::make_tuple(blk->coords[3 * w], blk->coords[3 * w + 1], blk->coords[3 * w + 2]); };
        std::vector<double> rt(E, 0.0);
        std::vector<std::vector<std::pair<int, int>>> collectives;
        std::vector<std::pair<int, int>> one_to_one;
        for (int i = 1; i <= n; ++i) {
            const int k = splits[i];
            if (k > 1) {
                std::vector<std::pair<int, int>> fdeps, bdeps;
                std::vector<std::vector<std::pair<int, int>>> sync;
                std::map<std::pair<int, int>, bool> seen;
                for (int j = 0; j < k; ++j) {
                    const int fs = fsubs[i][j], bs = bsubs[i][j];
                    for (int v : G.nodes[fs].succ) fdeps.push_back({fs, v});
                    for (int p : G.nodes[bs].pred) {
                        const auto& sc = G.nodes[bs].succ;
                        if (std::find(sc.begin(), sc.end(), p) != sc.end()) {             // bidirectional sync edge
                            if (!seen.count({p, bs}) && !seen.count({bs, p})) { sync.push_back({{p, bs}, {bs, p}}); seen[{p, bs}] = true; }
                        } else bdeps.push_back({p, bs});
                    }
                }
                for (auto* d : {&fdeps, &bdeps}) {
                    std::vector<std::tuple<int, int, int>> ps, cs;
                    for (auto& pr : *d) { ps.push_back(coord(pr.first)); cs.push_back(coord(pr.second)); }
                    std::sort(ps.begin(), ps.end()); std::sort(cs.begin(), cs.end());
                    if (ps == cs) collectives.push_back(*d); else one_to_one.insert(one_to_one.end(), d->begin(), d->end());
                }
                for (auto& s : sync) collectives.push_back(s);
            } else {
                for (int v : G.nodes[fwd_node[i]].succ) one_to_one.push_back({fwd_node[i], v});
                for (int p : G.nodes[bwd_node[i]].pred) one_to_one.push_back({p, bwd_node[i]});
            }
        }
        for (auto& col : collectives) {
            std::vector<int> cgs, racks, nds; std::vector<std::tuple<int, int, int>> servers;
            double message = 0.0;
            for (auto& pr : col) {
                for (int h : {pr.first, pr.second}) {
                    auto c = coord(h);
                    cgs.push_back(std::get<0>(c)); racks.push_back(std::get<1>(c)); nds.push_back(std::get<2>(c)); servers.push_back(c);
                }
                message += G.size[Graph::key(pr.first, pr.second)];
            }
            auto uniq = [](auto& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); return (int)v.size(); };
            const int n_servers = uniq(servers);
            const double t = n_servers == 1 ? 0.0 : all_reduce_time(message, uniq(nds), uniq(racks), uniq(cgs), blk->num_communication_groups,
                                                                    blk->channel_bandwidth, blk->latency, blk->io_latency);
            for (auto& pr : col) rt[dep_index[Graph::key(pr.first, pr.second)]] = t;
        }
        for (auto& pr : one_to_one) {                             // set_one_to_one_dep_run_time, applied last
            const int e = dep_index[Graph::key(pr.first, pr.second)];
            rt[e] = (coord(pr.first) == coord(pr.second) || sizes[e] == 0.0) ? 0.0 : one_to_one_base + sizes[e] / blk->channel_bandwidth;
        }
        for (int e = 0; e < E; ++e) { sched_cost[e] = rt[e]; run_time[e] = is_flow[e] ? rt[e] : 0.0; }      // RCE:542-560
    }

    // ---- channels: one per (src worker, dst worker) pair that carries a flow, numbered in sorted order ----
    std::vector<long> chan_key(E);
    std::vector<long> keys;
    for (int e = 0; e < E; ++e) {
        chan_key[e] = (long)op_worker[idx[deps[e].u]] * W + op_worker[dep_dst[e]];
        if (is_flow[e]) keys.push_back(chan_key[e]);
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::vector<uint16_t> dep_channel(E, RAMP_NO_CHANNEL);
    for (int e = 0; e < E; ++e)
        if (is_flow[e]) dep_channel[e] = (uint16_t)(std::lower_bound(keys.begin(), keys.end(), chan_key[e]) - keys.begin());

    // ---- SRPT priorities: highest cost -> 0 (stable) ----
    std::vector<int64_t> op_prio(N, 0), dep_prio(E, 0);
    for (int w = 0; w < W; ++w) {
        std::vector<int> members;
        for (int i = 0; i < N; ++i) if (op_worker[i] == w) members.push_back(i);
        std::stable_sort(members.begin(), members.end(), [&](int a, int b) { return op_cost[a] > op_cost[b]; });
        for (size_t r = 0; r < members.size(); ++r) op_prio[members[r]] = (int64_t)r;
    }
    if (reference_times) {                                       // ALL deps in graph edge order, stable descending by run time
        std::vector<int> order;
        for (int h : alive) for (int v : G.nodes[h].succ) order.push_back(dep_index[Graph::key(h, v)]);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sched_cost[a] > sched_cost[b]; });
        for (int r = 0; r < E; ++r) if (is_flow[order[r]]) dep_prio[order[r]] = r;
    } else {
        std::vector<int> fl;
        for (int e = 0; e < E; ++e) if (is_flow[e]) fl.push_back(e);
        std::stable_sort(fl.begin(), fl.end(), [&](int a, int b) { return run_time[a] > run_time[b]; });
        for (size_t r = 0; r < fl.size(); ++r) dep_prio[fl[r]] = (int64_t)r;
    }

    // ---- parents = predecessors that are not also successors (JOB:508-523) ----
    std::vector<uint16_t> n_parents(N, 0);
    for (int e = 0; e < E; ++e)
        if (!G.size.count(Graph::key(deps[e].v, deps[e].u))) n_parents[dep_dst[e]]++;

    memset(out, 0, sizeof(*out));
    out->n_ops = N; out->n_deps = E; out->n_workers = W; out->n_channels = (int32_t)keys.size();
    out->num_training_steps = num_training_steps; out->model_id = 0; out->degree = degree;
    out->op_cost = dup(op_cost); out->op_prio = dup(op_prio); out->op_worker = dup(op_worker); out->op_n_parents = dup(n_parents);
    out->row_ptr = dup(row_ptr); out->dep_dst = dup(dep_dst); out->dep_run_time = dup(run_time); out->dep_prio = dup(dep_prio);
    out->dep_channel = dup(dep_channel); out->dep_is_flow = dup(is_flow);
    if (aux) {
        std::vector<double> op_mem(N);
        std::vector<int32_t> order(N);
        for (int i = 0; i < N; ++i) { op_mem[i] = G.nodes[by_id[i]].mem; order[i] = idx[alive[i]]; }   // op indices in graph (dict) order
        aux->dep_size = dup(sizes); aux->op_mem = dup(op_mem); aux->node_order = dup(order);
    }
    return RAMP_OK;
}

// RampFirstFitOpPlacer (agents/placers/ramp_first_fit_op_placer.py:27-113 + agents/placers/utils.py:68-582) for one job on
// a cluster whose servers are described by free memory and a busy flag; the Python twin is ddls_b200/placer.py.
int ramp_first_fit_place(const ramp_forward_graph_t* g, const int32_t* splits, const ramp_cluster_state_t* st,
                         int32_t* server_out, int32_t* offset_out) {
    if (!g || !splits || !st || !server_out || !offset_out || g->n_fwd < 1) return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_first_fit_place: null argument or empty graph");
    const int n = g->n_fwd, C = st->shape[0], R = st->shape[1], S = st->shape[2];
    const int n_servers = C * R * S;
    auto sid = [&](int c, int r, int s) { return (c * R + r) * S + s; };
    std::vector<double> mem(st->free_mem, st->free_mem + n_servers);
    std::vector<std::vector<int>> parents(n + 1), children(n + 1);
    for (int e = 0; e < g->n_edges; ++e) {
        const int u = g->edge_src[e], v = g->edge_dst[e];
        if (u < 1 || u > n || v < 1 || v > n) return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_first_fit_place: an edge names an op outside 1..n_fwd");
        parents[v].push_back(u); children[u].push_back(v);
    }
    // topo_sort (utils.py:100-115)
    std::vector<std::vector<int>> left = parents;
    std::vector<int> sequence, queue;
    for (int v = 1; v <= n; ++v) if (left[v].empty()) { queue.push_back(v); sequence.push_back(v); }
    for (size_t q = 0; q < queue.size(); ++q) {
        const int v = queue[q];
        for (int c : children[v]) {
            erase_first(left[c], v);
            if (left[c].empty()) { queue.push_back(c); sequence.push_back(c); }
        }
    }
    std::vector<std::vector<int>> where(n + 1);
    auto check_block = [&](const std::vector<int>& block, double op_size) {     // utils.py:215-233
        if (block.empty()) return false;
        for (int s : block) {
            if (s < 0) return false;
            if (st->busy[s]) return false;
            if (mem[s] < op_size) return false;
        }
        return true;
    };
    auto get_block = [&](int bc, int br, int bs, int i, int j, int k) {          // utils.py:464-489
        std::vector<int> block;
        if (bs == -1) {
            for (int m = 0; m < bc; ++m) {
                const int c = (i + m) % (C + 1), r = (j + m) % (R + 1), s = k % S;
                block.push_back((c < C && r < R) ? sid(c, r, s) : -1);           // the reference would raise KeyError here
            }
        } else {
            for (int c = 0; c < bc; ++c) for (int r = 0; r < br; ++r) for (int s = 0; s < bs; ++s)
                block.push_back(sid((i + c) % C, (j + r) % R, (k + s) % S));
        }
        return block;
    };
    for (int op : sequence) {
        const int split = std::max(splits[op - 1], 1);
        const double need = g->act_size[op - 1] + g->par_size[op - 1];
        bool placed = false;
        for (int p : parents[op]) {                                               // parent_collective_placement, utils.py:258-314
            const std::vector<int>& sv = where[p];
            if ((int)sv.size() != split) continue;
            double avail = 0.0;
            for (int s : sv) avail += mem[s];
            if (avail >= need) {
                for (int s : sv) { mem[s] -= need / split; where[op].push_back(s); }
                placed = true;
                break;
            }
        }
        if (placed) continue;
        if (split > n_servers) return 1;                                          // regular_collective_placement, utils.py:333-383
        const double op_size = need / split;
        std::vector<std::tuple<int, int, int>> shapes;
        for (int i = 1; i <= split; ++i) {                                        // get_factor_pairs + get_block_shapes
            if (split % i) continue;
            const int p0 = split / i, p1 = i;
            const double var = std::sqrt((double)p0);
            if (std::fmod(var, 1.0) == 0.0 && var <= C && var <= R && p1 <= S) shapes.push_back({(int)var, (int)var, p1});
            if (p0 > C || p0 > R || p1 > S) continue;
            shapes.push_back({p0, 1, p1});
            shapes.push_back({p0, p1, 1});
        }
        shapes.push_back({split, split, -1});
        shapes.push_back({split, 1, 1});
        std::vector<int> block;
        bool found = false;
        for (auto& sh : shapes) {                                                 // ff_block, utils.py:394-443
            const int bc = std::get<0>(sh), br = std::get<1>(sh), bs = std::get<2>(sh);
            const int I = C - bc + 1, J = R - br + 1, K = S - bs + 1;
            if (I <= 0 || J <= 0 || K <= 0) continue;
            for (int i = 0; i < I && !found; ++i) for (int j = 0; j < J && !found; ++j) for (int k = 0; k < K && !found; ++k) {
                block = get_block(bc, br, bs, i, j, k);
                if (check_block(block, op_size)) found = true;
            }
            if (found) break;
        }
        if (!found) return 1;
        for (int s : block) { mem[s] -= op_size; where[op].push_back(s); }
    }
    int at = 0;
    for (int op = 1; op <= n; ++op) {
        offset_out[op - 1] = at;
        for (int s : where[op]) server_out[at++] = s;
    }
    offset_out[n] = at;
    return RAMP_OK;
}


// The same decision for MANY cluster states at once (the batched environment groups its episodes by occupancy pattern and asks
// once per distinct pattern): busy_words[k] is a bit set over the servers (bit i = server i is busy; free servers hold no job, so
// their worker's memory is empty: memory_capacity bytes free).  server_mask_out[k] receives the set of servers the job is put
// on, ok_out[k] = 1, or ok_out[k] = 0 when the job cannot be placed there.
int ramp_first_fit_place_many(const ramp_forward_graph_t* g, const int32_t* splits, const int32_t shape[3], double memory_capacity,
                              int32_t n_states, int32_t n_words, const uint64_t* busy_words, uint64_t* server_mask_out, uint8_t* ok_out) {
    if (!g || !splits || !shape || !busy_words || !server_mask_out || !ok_out || n_states < 0 || n_words < 1)
        return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_first_fit_place_many: null argument");
    const int n_servers = shape[0] * shape[1] * shape[2];
    if (n_servers > 64 * n_words) return ramp_internal_set_error(RAMP_ERR_BAD_ARG, "ramp_first_fit_place_many: n_words too small for the cluster");
    int total = 0;
    for (int i = 0; i < g->n_fwd; ++i) total += std::max(splits[i], 1);
    std::vector<double> free_mem(n_servers, memory_capacity);
    std::vector<uint8_t> busy(n_servers);
    std::vector<int32_t> server_out(total), offset_out(g->n_fwd + 1);
    ramp_cluster_state_t st{};
    st.shape[0] = shape[0]; st.shape[1] = shape[1]; st.shape[2] = shape[2];
    st.free_mem = free_mem.data(); st.busy = busy.data();
    for (int32_t k = 0; k < n_states; ++k) {
        const uint64_t* bw = busy_words + (size_t)k * n_words;
        for (int i = 0; i < n_servers; ++i) busy[i] = (uint8_t)((bw[i >> 6] >> (i & 63)) & 1ull);
        uint64_t* out = server_mask_out + (size_t)k * n_words;
        for (int w = 0; w < n_words; ++w) out[w] = 0ull;
        const int rc = ramp_first_fit_place(g, splits, &st, server_out.data(), offset_out.data());
        if (rc < 0) return rc;
        ok_out[k] = rc == RAMP_OK ? 1 : 0;
        if (rc == RAMP_OK) for (int i = 0; i < total; ++i) out[server_out[i] >> 6] |= 1ull << (server_out[i] & 63);
    }
    return RAMP_OK;
}

}  // extern "C"
