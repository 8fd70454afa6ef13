// allocator.cpp -- best-effort topology allocator.
//
// Same answers as internal/pkg/allocator/{device.go,besteffort_policy.go}: pair weights
// from kfd io_links/p2p_links files (or from a measured link list), grouping by physical
// GPU, two-phase candidate enumeration, first-strictly-smallest selection.  Built
// differently: link files are read once and parsed in a single pass, NodeIds are mapped to
// a dense index and the pair weights live in a flat matrix so the candidate search does
// no map lookups.
#include <algorithm>
#include <climits>
#include <cstring>
#include <unordered_map>
#include <unordered_set>

#include <atomic>
#include <thread>
#include "gosem.hpp"
#include "internal.hpp"

namespace b2dp {

// device.go:38-54
enum { kSameDevId = 10, kXgmiLink = 10, kSameNuma = 10, kDiffDevId = 20, kDiffNuma = 20, kPcieLink = 40, kOtherLink = 50 };

int calculate_pair_weight(const Device& a, const Device& b, int link_type) {
    int w = a.dev_id == b.dev_id ? kSameDevId : kDiffDevId;                       // device.go:137-141
    w += link_type == 11 ? kXgmiLink : link_type == 2 ? kPcieLink : kOtherLink;   // device.go:143-149
    w += a.numa == b.numa ? kSameNuma : kDiffNuma;                                // device.go:151-155
    return w;
}

struct Partitions {  // device.go:75-80
    std::string parent_id, dev_id;
    std::vector<int> ids;
};

struct DeviceSet {  // device.go:67-73
    std::vector<int> ids;
    std::vector<int> parents;
    int weight = 0;
};

class BestEffortPolicy {
public:
    std::vector<Device> devices;
    std::unordered_map<std::string, int> by_id;    // devicesMap
    std::map<std::string, Partitions> partitions;  // devicePartitions, keyed by DevId
    std::map<int, std::map<int, int>> p2p;         // p2pWeights[from][to], from < to
    std::unordered_map<int, int> node_index;       // dense view for the search
    std::vector<int> dense;                        // n x n, [lo][hi] only
    int n_nodes = 0;
    std::mutex mu;

    void build_dense() {
        node_index.clear();
        std::vector<int> nodes;
        for (auto& d : devices) nodes.push_back(d.node_id);
        for (auto& r : p2p) { nodes.push_back(r.first); for (auto& c : r.second) nodes.push_back(c.first); }
        std::sort(nodes.begin(), nodes.end());
        nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
        n_nodes = (int)nodes.size();
        for (int i = 0; i < n_nodes; ++i) node_index[nodes[i]] = i;
        dense.assign((size_t)n_nodes * n_nodes, 0);
        for (auto& r : p2p)
            for (auto& c : r.second) dense[(size_t)node_index[r.first] * n_nodes + node_index[c.first]] = c.second;
    }
    // p2pWeights[min(a,b)][max(a,b)], a missing entry reads as 0 (device.go:259-266)
    int weight(int node_a, int node_b) const {
        const int lo = std::min(node_a, node_b), hi = std::max(node_a, node_b);
        auto ia = node_index.find(lo), ib = node_index.find(hi);
        if (ia == node_index.end() || ib == node_index.end()) return 0;
        return dense[(size_t)ia->second * n_nodes + ib->second];
    }
};

BestEffortPolicy* policy_new() { return new BestEffortPolicy(); }
void policy_free(BestEffortPolicy* p) { delete p; }

// device.go:107-133, one pass: every line against every key, LAST match per key wins.
// Returns false where the reference returns an error (open failure, ParseInt(…, 0, 32) error).
static bool fetch_topo_properties(const std::string& path, const char* const* keys, int n_keys, int* res) {
    std::string data;
    for (int i = 0; i < n_keys; ++i) res[i] = 0;
    if (!go::read_attr(path, data)) return false;
    bool ok = true;
    go::scan_lines(data, [&](std::string_view line) {
        for (int i = 0; i < n_keys; ++i) {
            std::string_view digits;
            if (!go::match_key_digits(line, keys[i], digits)) continue;
            int64_t v;
            if (go::parse_int(digits, 0, 32, &v) != go::NumErr::none) { ok = false; return false; }
            res[i] = (int)v;
        }
        return true;
    });
    return ok;
}

// device.go:181-215 for one link record.
static void apply_link(const std::vector<Device>& devs, const std::unordered_set<int>& lookup, int node_from,
                       int node_to, int type, std::map<int, std::map<int, int>>& p2p) {
    const int from = node_from < node_to ? node_from : node_to;
    const int to = node_from < node_to ? node_to : node_from;
    if (!lookup.count(from) || !lookup.count(to)) return;
    const Device *fd = nullptr, *td = nullptr;
    bool found = false;
    for (const auto& d : devs) {  // device.go:198-209
        if (d.node_id == from) fd = &d;
        if (d.node_id == to) td = &d;
        if (fd && td) { found = true; break; }
    }
    if (found) p2p[from][to] = calculate_pair_weight(*fd, *td, type);
}

static void finish_init(BestEffortPolicy* p, const std::vector<Device>& devs) {
    // besteffort_policy.go:75-84 + device.go:287-304
    p->devices = devs;
    for (size_t i = 0; i < devs.size(); ++i) p->by_id[devs[i].id] = (int)i;
    p->partitions.clear();
    for (const auto& d : devs) {
        auto& ps = p->partitions[d.dev_id];
        ps.dev_id = d.dev_id;
        if (d.id.find("amdgpu_xcp") == std::string::npos) ps.parent_id = d.id;  // device.go:297
        ps.ids.push_back(d.node_id);
    }
    p->build_dense();
}

int policy_init_dir(BestEffortPolicy* p, const std::vector<Device>& devs, const std::string& dir_in) {
    std::lock_guard<std::mutex> g(p->mu);
    if (devs.empty()) return B2DP_E_ALLOC_EMPTY_DEVICES;  // device.go:221-225
    const std::string dir = dir_in.empty() ? "/sys/class/kfd/kfd/topology/nodes" : dir_in;  // device.go:226-228
    std::unordered_set<int> lookup;
    for (auto& d : devs) lookup.insert(d.node_id);
    static const char* const kMinor[] = {"drm_render_minor"};
    static const char* const kLink[] = {"node_from", "node_to", "type"};
    // The node directories are independent: each one's link files are listed and parsed (io_links before
    // p2p_links, each in Glob order) into that node's own record list, a few nodes at a time on a large
    // tree (CPX: 65 nodes, 4,034 link files, one open/read/close each).  The records are then applied
    // strictly in the reference's visiting order, because a later file overwrites an earlier one
    // (device.go:214).
    const std::vector<std::string> node_dirs = go::glob_digit_prefixed(dir);
    struct Rec { int v[3]; };
    std::vector<std::vector<Rec>> per_node(node_dirs.size());
    auto scan_node = [&](size_t i) {
        int minor;
        if (!fetch_topo_properties(node_dirs[i] + "/properties", kMinor, 1, &minor) || minor <= 0) return;  // device.go:240-244
        for (const char* sub : {"/io_links", "/p2p_links"})
            for (const auto& lp : go::glob_digit_prefixed(node_dirs[i] + sub)) {
                Rec r;
                if (fetch_topo_properties(lp + "/properties", kLink, 3, r.v)) per_node[i].push_back(r);  // device.go:176-179
            }
    };
    // measured on the 128-core GPU box host (CPX tree): 12.4 ms on 1 thread, 6.0 ms on 4, no gain beyond
    const size_t n_threads = node_dirs.size() >= 16 ? std::min<size_t>(4, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (n_threads <= 1) for (size_t i = 0; i < node_dirs.size(); ++i) scan_node(i);
    else {
        std::atomic<size_t> next{0};
        auto runner = [&] { for (size_t i; (i = next.fetch_add(1)) < node_dirs.size();) scan_node(i); };
        std::vector<std::thread> th;
        for (size_t t = 1; t < n_threads; ++t) th.emplace_back(runner);
        runner();
        for (auto& t : th) t.join();
    }
    for (const auto& recs : per_node)
        for (const Rec& r : recs) apply_link(devs, lookup, r.v[0], r.v[1], r.v[2], p->p2p);
    if (p->p2p.empty()) return B2DP_E_ALLOC_NO_WEIGHTS;  // besteffort_policy.go:72-74
    finish_init(p, devs);
    return B2DP_OK;
}

int policy_init_links(BestEffortPolicy* p, const std::vector<Device>& devs, const std::vector<Link>& links) {
    std::lock_guard<std::mutex> g(p->mu);
    if (devs.empty()) return B2DP_E_ALLOC_EMPTY_DEVICES;
    std::unordered_set<int> lookup;
    for (auto& d : devs) lookup.insert(d.node_id);
    for (const auto& l : links) apply_link(devs, lookup, l.from, l.to, l.type, p->p2p);
    if (p->p2p.empty()) return B2DP_E_ALLOC_NO_WEIGHTS;
    finish_init(p, devs);
    return B2DP_OK;
}

void policy_pair_weights(BestEffortPolicy* p, std::vector<b2dp_pair_weight>& out, int* n_rows) {
    std::lock_guard<std::mutex> g(p->mu);
    out.clear();
    for (auto& r : p->p2p)
        for (auto& c : r.second) out.push_back({r.first, c.first, c.second});
    *n_rows = (int)p->p2p.size();
}

int policy_group_count(BestEffortPolicy* p) {
    std::lock_guard<std::mutex> g(p->mu);
    return (int)p->partitions.size();
}

// device.go:254-273
static inline void add_device(const BestEffortPolicy* p, DeviceSet& s, int node) {
    int w = s.weight;
    for (int d : s.ids) w += p->weight(d, node);
    s.weight = w;
    s.ids.push_back(node);
}

// filterPartitions (device.go:310-351): one group per physical GPU holding the available,
// non-required NodeIds in ascending order; groups sorted by (len, ParentId), ties by DevId.
static std::vector<Partitions> build_groups(const BestEffortPolicy* p, const std::vector<int>& available,
                                            const std::vector<int>& required) {
    const auto& devs = p->devices;
    std::unordered_set<int> avail_nodes, req_nodes;
    for (int i : available) avail_nodes.insert(devs[i].node_id);
    for (int i : required) req_nodes.insert(devs[i].node_id);
    std::vector<Partitions> groups;
    for (const auto& kv : p->partitions) {
        Partitions f;
        for (int id : kv.second.ids)
            if (!req_nodes.count(id) && avail_nodes.count(id)) f.ids.push_back(id);
        if (f.ids.empty()) continue;
        std::sort(f.ids.begin(), f.ids.end());
        f.dev_id = kv.second.dev_id; f.parent_id = kv.second.parent_id;
        groups.push_back(std::move(f));
    }
    std::stable_sort(groups.begin(), groups.end(), [](const Partitions& a, const Partitions& b) {
        if (a.ids.size() != b.ids.size()) return a.ids.size() < b.ids.size();
        if (a.parent_id != b.parent_id) return a.parent_id < b.parent_id;
        return a.dev_id < b.dev_id;
    });
    return groups;
}

// device.go:353-442, literally: two-phase enumeration of ordered group sequences (phase 2 is a
// FIFO over partial sets).  Exponential in the number of groups; kept as the reference-shaped
// implementation (used by b2dp_allocator_candidates and as the fallback of the fast path).
static int candidate_subsets(const BestEffortPolicy* p, std::vector<int> available, const std::vector<int>& required,
                             int size, std::vector<DeviceSet>& finals) {
    finals.clear();
    if (size <= 0) return B2DP_E_ALLOC_SUBSET_SIZE;
    if ((int)available.size() < size) return B2DP_E_ALLOC_SUBSET_AVAIL;
    const auto& devs = p->devices;
    std::vector<Partitions> groups = build_groups(p, available, required);

    const int new_size = size - (int)required.size();
    const int n_groups = (int)groups.size();
    auto add_required = [&](DeviceSet& s) { for (int r : required) add_device(p, s, devs[r].node_id); };

    std::vector<DeviceSet> temp;
    for (int idx = 0; idx < n_groups; ++idx) {  // phase 1, device.go:375-402
        const auto& part = groups[idx];
        DeviceSet s;
        s.ids.push_back(part.ids[0]);
        s.parents.push_back(idx);
        if (new_size == 1) { add_required(s); finals.push_back(std::move(s)); continue; }
        bool fulfilled = false;
        for (int i = 1; i < (int)part.ids.size(); ++i) {
            add_device(p, s, part.ids[i]);
            if (i == new_size - 1) { fulfilled = true; break; }
        }
        if (fulfilled) { add_required(s); finals.push_back(std::move(s)); }
        else temp.push_back(std::move(s));
    }
    for (size_t head = 0; head < temp.size(); ++head) {  // phase 2 (FIFO), device.go:405-440
        // copy: temp may reallocate while we push
        const DeviceSet cur = temp[head];
        if ((int)cur.parents.size() == n_groups) continue;
        for (int idx = 0; idx < n_groups; ++idx) {
            if (std::find(cur.parents.begin(), cur.parents.end(), idx) != cur.parents.end()) continue;
            DeviceSet s = cur;
            s.parents.push_back(idx);
            bool done = false;
            for (int id : groups[idx].ids) {
                add_device(p, s, id);
                if ((int)s.ids.size() == new_size) {
                    add_required(s);
                    finals.push_back(s);
                    done = true;
                    break;
                }
            }
            if (!done && (int)s.ids.size() < new_size) temp.push_back(std::move(s));
        }
    }
    return B2DP_OK;
}

// ---- exact fast path -----------------------------------------------------------------------
// The reference enumerates ORDERED sequences of groups: every group but the last is taken whole,
// the last contributes its lowest NodeIds until the size is reached; finals appear level by level
// (number of groups used) and, within a level, in lexicographic order of the sequence; the first
// strictly smallest total weight wins.  The weight of a set does not depend on the order its members
// were added, so all orderings of the same (whole groups S, last group l) tie and the earliest of
// them is sorted(S) followed by l.  The winner is therefore
//      argmin over valid (S, l) of (weight, |S|+1, sorted(S)+[l])           -- 2^G * G states
// instead of P(G, k) sequences (8 GPUs, size 7: 1,024 states instead of 40,320 sequences), with a
// subset DP for the weights.  Same answer as the FIFO enumeration, bit for bit, including ties.
constexpr int kFastMaxGroups = 16;

static bool best_candidate_fast(const BestEffortPolicy* p, const std::vector<Partitions>& groups,
                                const std::vector<int>& req_nodes, int new_size, DeviceSet& best_set, bool& found,
                                double& n_sequences) {
    const int G = (int)groups.size();
    found = false;
    n_sequences = 0;
    if (G > kFastMaxGroups || new_size < 1) return false;  // not applicable: caller falls back
    if (G == 0) return true;
    // per-group prefix tables
    std::vector<std::vector<int>> intra(G), reqp(G);            // [g][c]: first c ids of g among themselves / vs required
    std::vector<std::vector<std::vector<int>>> crossp(G, std::vector<std::vector<int>>(G));  // [l][h][c]: prefix c of l vs all of h
    int req_intra = 0;
    for (size_t a = 0; a < req_nodes.size(); ++a)
        for (size_t b = 0; b < a; ++b) req_intra += p->weight(req_nodes[b], req_nodes[a]);
    for (int g = 0; g < G; ++g) {
        const auto& ids = groups[g].ids;
        intra[g].assign(ids.size() + 1, 0);
        reqp[g].assign(ids.size() + 1, 0);
        for (size_t c = 0; c < ids.size(); ++c) {
            int add = 0, radd = 0;
            for (size_t j = 0; j < c; ++j) add += p->weight(ids[j], ids[c]);
            for (int r : req_nodes) radd += p->weight(ids[c], r);
            intra[g][c + 1] = intra[g][c] + add;
            reqp[g][c + 1] = reqp[g][c] + radd;
        }
        for (int h = 0; h < G; ++h) {
            if (h == g) continue;
            crossp[g][h].assign(ids.size() + 1, 0);
            for (size_t c = 0; c < ids.size(); ++c) {
                int add = 0;
                for (int y : groups[h].ids) add += p->weight(ids[c], y);
                crossp[g][h][c + 1] = crossp[g][h][c] + add;
            }
        }
    }
    auto full = [&](int g) { return (int)groups[g].ids.size(); };
    // subset DP over whole groups: total size and weight (incl. pairs with the required devices)
    const uint32_t n_sub = 1u << G;
    std::vector<int> ssize(n_sub, 0), sweight(n_sub, 0);
    for (uint32_t S = 1; S < n_sub; ++S) {
        const int g = __builtin_ctz(S);
        const uint32_t R = S & (S - 1);
        int w = sweight[R] + intra[g][full(g)] + reqp[g][full(g)];
        for (uint32_t T = R; T; T &= T - 1) { const int h = __builtin_ctz(T); w += crossp[g][h][full(g)]; }
        sweight[S] = w;
        ssize[S] = ssize[R] + full(g);
    }
    std::vector<double> fact(G + 1, 1.0);
    for (int i = 1; i <= G; ++i) fact[i] = fact[i - 1] * i;
    // (S, l) precedes (S2, l2) in the reference's order of equal-weight finals?
    auto seq_less = [&](uint32_t S, int l, uint32_t S2, int l2) {
        const int k = __builtin_popcount(S), k2 = __builtin_popcount(S2);
        if (k != k2) return k < k2;
        uint32_t a = S, b = S2;
        while (a && b) {
            const int x = __builtin_ctz(a), y = __builtin_ctz(b);
            if (x != y) return x < y;
            a &= a - 1; b &= b - 1;
        }
        return l < l2;
    };
    int best_w = INT32_MAX, best_l = -1;
    uint32_t best_S = 0;
    for (uint32_t S = 0; S < n_sub; ++S) {
        if (ssize[S] >= new_size) continue;  // every whole-group prefix must still be short
        const int need = new_size - ssize[S];
        const int k = __builtin_popcount(S);
        for (int l = 0; l < G; ++l) {
            if (S & (1u << l)) continue;
            if (full(l) < need) continue;
            int w = sweight[S] + intra[l][need] + reqp[l][need] + req_intra;
            for (uint32_t T = S; T; T &= T - 1) w += crossp[l][__builtin_ctz(T)][need];
            n_sequences += fact[k];
            if (w < best_w || (w == best_w && found && seq_less(S, l, best_S, best_l))) {
                best_w = w; best_S = S; best_l = l; found = true;
            }
        }
    }
    if (!found) return true;
    best_set.ids.clear();
    best_set.weight = best_w;
    for (uint32_t T = best_S; T; T &= T - 1)
        for (int id : groups[__builtin_ctz(T)].ids) best_set.ids.push_back(id);
    const int need = new_size - ssize[best_S];
    for (int c = 0; c < need; ++c) best_set.ids.push_back(groups[best_l].ids[c]);
    for (int r : req_nodes) best_set.ids.push_back(r);
    return true;
}

int policy_allocate(BestEffortPolicy* p, const std::vector<std::string>& avail, const std::vector<std::string>& req,
                    int size, std::vector<std::string>& out, int* n_candidates, int* best_weight,
                    bool candidates_only) {
    std::lock_guard<std::mutex> g(p->mu);
    out.clear();
    if (n_candidates) *n_candidates = 0;
    if (best_weight) *best_weight = 0;
    auto to_indices = [&](const std::vector<std::string>& ids, std::vector<int>& idx) {
        for (const auto& s : ids) {
            auto it = p->by_id.find(s);
            if (it == p->by_id.end()) return false;  // nil *Device => the reference panics
            idx.push_back(it->second);
        }
        return true;
    };
    if (!candidates_only) {  // besteffort_policy.go:88-124
        if (size <= 0) return B2DP_E_ALLOC_SIZE;
        if ((int)avail.size() < size) return B2DP_E_ALLOC_AVAILABLE;
        if ((int)req.size() > size) return B2DP_E_ALLOC_REQUIRED;
        if (req.size() > avail.size()) return B2DP_E_ALLOC_REQ_AVAILABLE;
        if (p->devices.empty()) return B2DP_E_ALLOC_INIT;
        if ((int)avail.size() == size) { out = avail; return B2DP_OK; }
        if ((int)req.size() == size) { out = req; return B2DP_OK; }
        if (p->p2p.empty()) return B2DP_E_ALLOC_INIT;
        for (const auto& r : req)  // setContainsAll, device.go:88-105
            if (std::find(avail.begin(), avail.end(), r) == avail.end()) return B2DP_E_ALLOC_NOCANDIDATE;
    }
    std::vector<int> a_idx, r_idx;
    if (!to_indices(avail, a_idx) || !to_indices(req, r_idx)) return B2DP_E_PANIC;
    // device.go:362-364: available sorted by NodeId (stable here; Go's sort.Slice is not, which
    // only matters for duplicate NodeIds)
    std::stable_sort(a_idx.begin(), a_idx.end(),
                     [&](int x, int y) { return p->devices[x].node_id < p->devices[y].node_id; });
    if (!candidates_only) {
        // exact fast path (see best_candidate_fast); identical result to the enumeration below
        if (size <= 0) return B2DP_E_ALLOC_SUBSET_SIZE;
        if ((int)a_idx.size() < size) return B2DP_E_ALLOC_SUBSET_AVAIL;
        std::vector<int> req_nodes;
        for (int r : r_idx) req_nodes.push_back(p->devices[r].node_id);
        DeviceSet bs;
        bool found = false;
        double nseq = 0;
        if (best_candidate_fast(p, build_groups(p, a_idx, r_idx), req_nodes, size - (int)r_idx.size(), bs, found, nseq)) {
            if (n_candidates) *n_candidates = nseq > 2147483647.0 ? INT32_MAX : (int)nseq;
            if (best_weight) *best_weight = found ? bs.weight : 0;
            if (!found) return B2DP_E_PANIC;  // nil candidate deref, besteffort_policy.go:141
            for (int id : bs.ids)             // besteffort_policy.go:141-148
                for (int ai : a_idx)
                    if (p->devices[ai].node_id == id) { out.push_back(p->devices[ai].id); break; }
            return B2DP_OK;
        }
    }
    std::vector<DeviceSet> finals;
    int rc = candidate_subsets(p, a_idx, r_idx, size, finals);
    if (rc != B2DP_OK) return rc;
    int best = INT32_MAX;  // besteffort_policy.go:133-140
    const DeviceSet* cand = nullptr;
    for (const auto& s : finals)
        if (s.weight < best) { cand = &s; best = s.weight; }
    if (n_candidates) *n_candidates = (int)finals.size();
    if (best_weight) *best_weight = cand ? cand->weight : 0;
    if (candidates_only) return B2DP_OK;
    if (!cand) return B2DP_E_PANIC;  // nil candidate deref, besteffort_policy.go:141
    for (int id : cand->ids)         // besteffort_policy.go:141-148
        for (int ai : a_idx)
            if (p->devices[ai].node_id == id) { out.push_back(p->devices[ai].id); break; }
    return B2DP_OK;
}

}  // namespace b2dp

// ================================ C ABI ====================================================
using namespace b2dp;

struct b2dp_allocator { BestEffortPolicy* p; };

extern "C" int b2dp_allocator_new(b2dp_allocator** out) {
    if (!out) return B2DP_E_INVAL;
    *out = new b2dp_allocator{policy_new()};
    return B2DP_OK;
}
extern "C" void b2dp_allocator_free(b2dp_allocator* a) {
    if (!a) return;
    policy_free(a->p);
    delete a;
}
static std::vector<Device> devs_from_abi(const b2dp_device* devs, int n) {
    std::vector<Device> v;
    for (int i = 0; i < n; ++i) v.push_back(from_abi(devs[i]));
    return v;
}
extern "C" int b2dp_allocator_init(b2dp_allocator* a, const b2dp_device* devs, int n, const char* topo_nodes_dir) {
    if (!a || n < 0 || (n && !devs)) return B2DP_E_INVAL;
    return policy_init_dir(a->p, devs_from_abi(devs, n), topo_nodes_dir ? topo_nodes_dir : "");
}
extern "C" int b2dp_allocator_init_links(b2dp_allocator* a, const b2dp_device* devs, int n, const b2dp_link* links,
                                         int n_links) {
    if (!a || n < 0 || (n && !devs) || n_links < 0 || (n_links && !links)) return B2DP_E_INVAL;
    std::vector<Link> l;
    for (int i = 0; i < n_links; ++i) l.push_back({links[i].node_from, links[i].node_to, links[i].type});
    return policy_init_links(a->p, devs_from_abi(devs, n), l);
}
extern "C" int b2dp_allocator_pair_weights(b2dp_allocator* a, b2dp_pair_weight* out, int cap, int* n, int* n_rows) {
    if (!a || !n || cap < 0) return B2DP_E_INVAL;
    std::vector<b2dp_pair_weight> v;
    int rows = 0;
    policy_pair_weights(a->p, v, &rows);
    *n = (int)v.size();
    if (n_rows) *n_rows = rows;
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && !out) return B2DP_E_INVAL;
    memcpy(out, v.data(), v.size() * sizeof(b2dp_pair_weight));
    return B2DP_OK;
}
extern "C" int b2dp_allocator_group_count(b2dp_allocator* a, int32_t* groups) {
    if (!a || !groups) return B2DP_E_INVAL;
    *groups = policy_group_count(a->p);
    return B2DP_OK;
}
static std::vector<std::string> strs(const char* const* p, int n) {
    std::vector<std::string> v;
    for (int i = 0; i < n; ++i) v.emplace_back(p[i] ? p[i] : "");
    return v;
}
extern "C" int b2dp_allocator_candidates(b2dp_allocator* a, const char* const* available, int na,
                                         const char* const* required, int nr, int size, int32_t* n_candidates,
                                         int32_t* best_weight) {
    if (!a || na < 0 || nr < 0 || (na && !available) || (nr && !required)) return B2DP_E_INVAL;
    std::vector<std::string> out;
    int nc = 0, bw = 0;
    int rc = policy_allocate(a->p, strs(available, na), strs(required, nr), size, out, &nc, &bw, true);
    if (n_candidates) *n_candidates = nc;
    if (best_weight) *best_weight = bw;
    return rc;
}
extern "C" int b2dp_allocator_allocate(b2dp_allocator* a, const char* const* available, int na,
                                       const char* const* required, int nr, int size, char (*out)[64], int cap,
                                       int* n) {
    if (!a || !n || na < 0 || nr < 0 || (na && !available) || (nr && !required) || cap < 0) return B2DP_E_INVAL;
    std::vector<std::string> ids;
    *n = 0;
    int rc = policy_allocate(a->p, strs(available, na), strs(required, nr), size, ids, nullptr, nullptr, false);
    if (rc != B2DP_OK) return rc;
    *n = (int)ids.size();
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && !out) return B2DP_E_INVAL;
    for (int i = 0; i < *n; ++i) copy_str(out[i], 64, ids[i]);
    return B2DP_OK;
}
