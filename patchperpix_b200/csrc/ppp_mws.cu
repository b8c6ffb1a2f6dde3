// Mutex-watershed partition of the patch graph -- the flylight default
// (`mws = true`), reference: PatchPerPix/vote_instances/graph_mws.py:7-85 on the
// graph that setAffgraph builds (aff_patch_graph.py:31-40).
//
// HOST code by design.  The algorithm is one greedy pass over the edges in
// order of decreasing |aff| where every decision depends on all earlier ones;
// the graph of one block has a few thousand edges (well under a millisecond on one
// core, less than a single dependent-load chain would cost on the device); the global
// graph of a blockwise run has ~10^6 (FlyLight-sized volume: 0.98 M edges), which is why
// the graph is built with flat tables, the edge order comes from a counting sort, the
// edges are radix-sorted and the exclusions live in one flat set (1 M edges: 0.18 s on a
// slow core, was 1.4 s with node-based hash maps and a comparison sort).  Everything
// around it stays on the GPU: the affinities come from ppp_patch_graph, the labels go to
// ppp_paint.
//
// What has to be reproduced exactly, because label VALUES depend on it:
//  * edge order = networkx edge iteration (nodes in insertion order, neighbours
//    in insertion order, each edge once), then a stable sort by |aff|, descending;
//  * component ids: a new component takes max(id in use) + 1, a merge keeps the
//    smaller id, ids of merged-away components are skipped in the numbering
//    unless they were the maximum (then they are taken again);
//  * nodes that never join through an accepted attractive edge get no label;
//  * the mutex test is on clusters (a cluster-level restatement of the
//    reference's scan over all mutex edges, same outcome).
#include "../../include/ppp_b200.h"
#include "ppp_api.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct MwsEdge { int u, v; float aff; };             // aff signed: > 0 attractive

inline uint32_t abs_bits(float a)
{
    uint32_t k;
    memcpy(&k, &a, 4);
    return k & 0x7fffffffu;
}

// open addressing, u64 key -> int value, sized once (keys are never removed)
struct FlatMap {
    std::vector<uint64_t> key;
    std::vector<int> val;
    size_t mask;
    explicit FlatMap(int64_t expected)
    {
        size_t cap = 16;
        while ((int64_t)cap < 2 * expected + 16) cap <<= 1;
        if (expected <= 0) cap = 1;
        key.assign(cap, ~0ULL);
        val.assign(cap, 0);
        mask = cap - 1;
    }
    // value of `k`; inserted with `fresh_val` if absent (*fresh says which)
    int& at(uint64_t k, int fresh_val, bool* fresh)
    {
        uint64_t h = k * 0x9E3779B97F4A7C15ULL;
        size_t i = (size_t)(h >> 20) & mask;
        while (true) {
            if (key[i] == k) { *fresh = false; return val[i]; }
            if (key[i] == ~0ULL) { key[i] = k; val[i] = fresh_val; *fresh = true; return val[i]; }
            i = (i + 1) & mask;
        }
    }
};

// growing set of u64 keys (open addressing, half full at most)
struct FlatSet {
    std::vector<uint64_t> key;
    size_t mask, used;
    FlatSet() : key(1024, ~0ULL), mask(1023), used(0) {}
    static size_t home(uint64_t k, size_t mask)
    {
        return (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & mask;
    }
    bool contains(uint64_t k) const
    {
        for (size_t i = home(k, mask);; i = (i + 1) & mask) {
            if (key[i] == k) return true;
            if (key[i] == ~0ULL) return false;
        }
    }
    bool insert(uint64_t k)                              // false: was there already
    {
        for (size_t i = home(k, mask);; i = (i + 1) & mask) {
            if (key[i] == k) return false;
            if (key[i] == ~0ULL) { key[i] = k; break; }
        }
        if (++used * 2 > mask) {
            std::vector<uint64_t> old;
            old.swap(key);
            mask = mask * 2 + 1;
            key.assign(mask + 1, ~0ULL);
            for (uint64_t q : old)
                if (q != ~0ULL) {
                    size_t i = home(q, mask);
                    while (key[i] != ~0ULL) i = (i + 1) & mask;
                    key[i] = q;
                }
        }
        return true;
    }
};

// clusters with their mutual exclusions.  The relation "root a must not join root b" is
// one set of root pairs; every root also keeps the list of partners it was ever excluded
// from (ids that may have been merged away since: find() gives their present root), so
// that a join can re-key the pairs of the root that disappears.  Pairs of a vanished root
// stay in the set: a non-root is never asked about again.
struct Clusters {
    std::vector<int> parent;
    std::vector<std::vector<int>> partners;
    FlatSet pairs;
    explicit Clusters(int n) : parent(n), partners(n) { for (int i = 0; i < n; i++) parent[i] = i; }
    static uint64_t pair_key(int a, int b)
    {
        return ((uint64_t)(uint32_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
    }
    int find(int a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    }
    bool exclusive(int a, int b) const { return pairs.used != 0 && pairs.contains(pair_key(a, b)); }
    void forbid(int a, int b) {
        if (a == b || !pairs.insert(pair_key(a, b))) return;
        partners[a].push_back(b);
        partners[b].push_back(a);
    }
    int join(int a, int b) {                          // returns the surviving root
        if (a == b) return a;
        if (partners[a].size() < partners[b].size()) std::swap(a, b);
        parent[b] = a;
        for (int x : partners[b]) {
            const int c = find(x);
            if (c != a && pairs.insert(pair_key(a, c))) partners[a].push_back(c);
        }
        std::vector<int>().swap(partners[b]);
        return a;
    }
};

}  // namespace

extern "C" int ppp_mws_host(const uint32_t* pairs, const float* aff, int64_t n,
                            const ppp_cfg* cfg, int32_t* node_vox, int32_t* node_label,
                            int64_t* n_nodes, int32_t* n_labels)
{
    if (!cfg || !n_nodes || !n_labels || (n > 0 && (!pairs || !aff || !node_vox || !node_label)))
        return ppp_fail(-1, "ppp_mws_host: null argument");
    const int64_t Y = cfg->Y, X = cfg->X;

    const bool prof = getenv("PPP_MWS_PROF") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lapse = [&](const char* what) {
        if (!prof) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mws] %-10s %.1f ms\n", what,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    // --- the graph, with networkx's iteration order --------------------------
    // node ids in order of first appearance; flat tables instead of node-based hash maps
    // (the global graph of a blockwise run has ~10^6 edges and this pass is replicated on
    // every rank): a direct array when the voxel space is small (the blockwise driver hands
    // in compacted node ids), else open addressing
    const int64_t V = (int64_t)cfg->Z * Y * X;
    const bool direct = V > 0 && V <= (int64_t)(8 * n + 1024);
    std::vector<int> id_direct;
    FlatMap id_hash(direct ? 0 : 2 * n);
    if (direct) id_direct.assign((size_t)V, -1);
    std::vector<int64_t> vox;
    auto node = [&](const uint32_t* c) {
        const int64_t v = ((int64_t)c[0] * Y + c[1]) * X + c[2];
        if (direct && v < V) {
            int& id = id_direct[(size_t)v];
            if (id < 0) { id = (int)vox.size(); vox.push_back(v); }
            return id;
        }
        bool fresh;
        int& id = id_hash.at((uint64_t)v, (int)vox.size(), &fresh);
        if (fresh) vox.push_back(v);
        return id;
    };
    // the rows with a non-zero affinity, as (earlier node, later node, aff)
    std::vector<int> rlo, rhi;
    std::vector<float> raff;
    rlo.reserve(n); rhi.reserve(n); raff.reserve(n);
    for (int64_t i = 0; i < n; i++) {
        if (aff[i] == 0.0f) continue;                     // aff_patch_graph.py:36
        const int u = node(pairs + 6 * i), v = node(pairs + 6 * i + 3);
        rlo.push_back(std::min(u, v));
        rhi.push_back(std::max(u, v));
        raff.push_back(aff[i]);
    }
    lapse("graph");
    const int nn = (int)vox.size();
    const int64_t nr = (int64_t)raff.size();
    // networkx yields, for every node in insertion order, its neighbours in insertion order,
    // each edge once: from the EARLIER of its two nodes, at the place the pair was first
    // added to that node's neighbour list -- the order of the rows whose earlier node it is.
    // Stable counting sort by the earlier node; a pair that comes again keeps its place and
    // takes the later affinity (the attribute is overwritten).
    std::vector<int64_t> start((size_t)nn + 1, 0);
    for (int64_t i = 0; i < nr; i++) start[rlo[i] + 1]++;
    for (int i = 0; i < nn; i++) start[i + 1] += start[i];
    std::vector<int> s_hi((size_t)nr);
    std::vector<float> s_aff((size_t)nr);
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < nr; i++) {
            const int64_t q = fill[rlo[i]]++;
            s_hi[q] = rhi[i];
            s_aff[q] = raff[i];
        }
    }
    std::vector<MwsEdge> edges;
    edges.reserve((size_t)nr);
    {
        std::vector<int> stamp((size_t)nn, -1), where((size_t)nn, 0);
        for (int u = 0; u < nn; u++)
            for (int64_t q = start[u]; q < start[u + 1]; q++) {
                const int v = s_hi[q];
                if (stamp[v] != u) {
                    stamp[v] = u;
                    where[v] = (int)edges.size();
                    edges.push_back({u, v, s_aff[q]});    // graph_mws.py:22-26
                } else {
                    edges[where[v]].aff = s_aff[q];
                }
            }
    }
    lapse("edges");
    // stable sort by |aff|, descending (:28): LSD radix sort on the float bits of |aff|
    // (non-negative floats order like their bit patterns; the key is complemented),
    // 11 + 10 + 10 bits
    {
        std::vector<MwsEdge> tmp(edges.size());
        std::vector<MwsEdge>* src = &edges;
        std::vector<MwsEdge>* dst = &tmp;
        const int shift[3] = {0, 11, 21}, bits[3] = {11, 10, 10};
        for (int pass = 0; pass < 3; pass++) {
            const uint32_t dm = (1u << bits[pass]) - 1u;
            std::vector<size_t> hist((size_t)dm + 2, 0);
            for (const MwsEdge& e : *src) hist[(((~abs_bits(e.aff)) >> shift[pass]) & dm) + 1]++;
            for (uint32_t i = 0; i <= dm; i++) hist[i + 1] += hist[i];
            for (const MwsEdge& e : *src)
                (*dst)[hist[((~abs_bits(e.aff)) >> shift[pass]) & dm]++] = e;
            std::swap(src, dst);
        }
        edges.swap(tmp);                                  // three passes: the result is in tmp
    }
    lapse("sort");

    // --- the greedy pass (:33-75) --------------------------------------------
    Clusters cl(nn);
    std::vector<int> cid(nn, 0);          // by root: the reference's component id, 0 = none yet
    std::vector<int> members(1, 0);       // by component id: nodes carrying it
    int max_id = 0;                       // np.max(node_CCs.values())
    for (const MwsEdge& e : edges) {
        int a = cl.find(e.u), b = cl.find(e.v);
        if (!(e.aff > 0)) { cl.forbid(a, b); continue; }
        int ca = cid[a], cb = cid[b];
        if (ca == 0 && cb == 0) {                         // :37-43 (no mutex test here)
            int id = max_id + 1;
            if ((int)members.size() <= id) members.resize(id + 1, 0);
            members[id] = (e.u == e.v) ? 1 : 2;
            cid[cl.join(a, b)] = id;
            max_id = id;
        } else if (ca == 0 || cb == 0) {                  // :45-57
            if (cl.exclusive(a, b)) continue;
            int id = std::max(ca, cb);
            cid[cl.join(a, b)] = id;
            members[id] += 1;
        } else if (ca != cb) {                            // :58-73
            if (cl.exclusive(a, b)) continue;
            int keep = std::min(ca, cb), gone = std::max(ca, cb);
            cid[cl.join(a, b)] = keep;
            members[keep] += members[gone];
            members[gone] = 0;
            while (max_id > 0 && members[max_id] == 0) max_id--;
        }
    }
    lapse("greedy");
    for (int i = 0; i < nn; i++) {
        node_vox[i] = (int32_t)vox[i];
        node_label[i] = cid[cl.find(i)];
    }
    *n_nodes = nn;
    // ids ever created = length of the reference's component list (graph_mws.py:78-81);
    // ids merged away stay in it as empty components
    *n_labels = (int32_t)members.size() - 1;
    return 0;
}

// ---------------------------------------------------------------------------
// Order in which CPython iterates over set((i0,j0), (i1,j1), ...) built by
// inserting the pairs one by one.  The reference enumerates its patch pairs by
// iterating the python set that scipy's cKDTree.query_pairs returns
// (aff_patch_graph.py:57-110, stitch_patch_graph.py:236-258), and that order
// decides the numbering of the instances.  Building the set of tuples and
// turning it back into an array dominates the host side of a block (3 of 8 ms on
// a 98^3 block), so the table is replayed here instead: tuple hash (xxHash-style,
// Objects/tupleobject.c, CPython >= 3.8), open addressing with 9 linear probes
// and a perturbed jump, growth x4 at 60 % fill (Objects/setobject.c).  The python
// side checks this against a real set once per process and falls back to it on
// any difference (other interpreter version).
// pairs i64 [n][2] (non-negative, distinct); order i64 [n] out: order[k] = index
// of the k-th pair the set yields.
// ---------------------------------------------------------------------------
namespace {

inline uint64_t py_tuple2_hash(uint64_t a, uint64_t b)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL,
                   P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    const uint64_t lanes[2] = {a, b};
    for (int i = 0; i < 2; i++) {
        acc += lanes[i] * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += 2ULL ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t)-1) return 1546275796ULL;
    return acc;
}

struct PySetReplay {
    std::vector<int64_t> slot;        // index of the pair in a table slot, -1 = unused
    const std::vector<uint64_t>& hash;
    size_t mask, fill;
    explicit PySetReplay(const std::vector<uint64_t>& h) : slot(8, -1), hash(h), mask(7), fill(0) {}
    static void insert_clean(std::vector<int64_t>& t, size_t mask, int64_t key, uint64_t h)
    {
        size_t perturb = h, i = (size_t)h & mask;
        while (true) {
            if (t[i] < 0) { t[i] = key; return; }
            if (i + 9 <= mask)
                for (size_t j = 1; j <= 9; j++)
                    if (t[i + j] < 0) { t[i + j] = key; return; }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void add(int64_t key)
    {
        const uint64_t h = hash[key];
        size_t perturb = h, i = (size_t)h & mask;
        while (true) {
            const size_t probes = (i + 9 <= mask) ? 9 : 0;
            bool placed = false;
            for (size_t j = 0; j <= probes; j++)
                if (slot[i + j] < 0) { slot[i + j] = key; placed = true; break; }
            if (placed) break;
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
        fill++;
        if (fill * 5 < mask * 3) return;
        const size_t minused = fill > 50000 ? fill * 2 : fill * 4;
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<int64_t> t(newsize, -1);
        for (size_t s = 0; s <= mask; s++)
            if (slot[s] >= 0) insert_clean(t, newsize - 1, slot[s], hash[slot[s]]);
        slot.swap(t);
        mask = newsize - 1;
    }
};

}  // namespace

extern "C" int ppp_pyset_order(const int64_t* pairs, int64_t n, int64_t* order)
{
    if (n < 0 || (n > 0 && (!pairs || !order))) return ppp_fail(-1, "ppp_pyset_order: null argument");
    std::vector<uint64_t> h((size_t)n);
    for (int64_t k = 0; k < n; k++) {
        if (pairs[2 * k] < 0 || pairs[2 * k + 1] < 0)
            return ppp_fail(-1, "ppp_pyset_order: negative index");
        h[k] = py_tuple2_hash((uint64_t)pairs[2 * k], (uint64_t)pairs[2 * k + 1]);
    }
    PySetReplay set(h);
    for (int64_t k = 0; k < n; k++) set.add(k);
    int64_t out = 0;
    for (size_t s = 0; s <= set.mask; s++)
        if (set.slot[s] >= 0) order[out++] = set.slot[s];
    return out == n ? 0 : ppp_fail(-1, "ppp_pyset_order: internal error");
}

// same replay, followed by the reference's distance filter (aff_patch_graph.py:61-69:
// a pair is dropped if |delta_d| > thr[d] on any axis), written out in set order:
// out i64 [<= n][2], *n_out pairs kept.  pts u32 [m][3] (the points the indices refer to).
extern "C" int ppp_pyset_pairs(const int64_t* pairs, int64_t n, const uint32_t* pts,
                               const double* thr, int64_t* out, int64_t* n_out)
{
    if (!n_out || n < 0 || (n > 0 && (!pairs || !pts || !thr || !out)))
        return ppp_fail(-1, "ppp_pyset_pairs: null argument");
    std::vector<uint64_t> h((size_t)n);
    for (int64_t k = 0; k < n; k++) {
        if (pairs[2 * k] < 0 || pairs[2 * k + 1] < 0)
            return ppp_fail(-1, "ppp_pyset_pairs: negative index");
        h[k] = py_tuple2_hash((uint64_t)pairs[2 * k], (uint64_t)pairs[2 * k + 1]);
    }
    PySetReplay set(h);
    for (int64_t k = 0; k < n; k++) set.add(k);
    int64_t kept = 0;
    for (size_t s = 0; s <= set.mask; s++) {
        const int64_t k = set.slot[s];
        if (k < 0) continue;
        const int64_t i = pairs[2 * k], j = pairs[2 * k + 1];
        bool drop = false;
        for (int d = 0; d < 3; d++) {
            const double delta = (double)pts[3 * i + d] - (double)pts[3 * j + d];
            if ((delta < 0 ? -delta : delta) > thr[d]) drop = true;
        }
        if (drop) continue;
        out[2 * kept] = i;
        out[2 * kept + 1] = j;
        kept++;
    }
    *n_out = kept;
    return 0;
}
