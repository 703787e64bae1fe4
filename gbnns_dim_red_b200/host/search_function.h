// search_function.h — source-level drop-in for the reference's search/search_function.h (and the
// support_func.h / support_classes.h / visited_list_pool.h it pulls in), backed by the B200 library.
//
// The reference's drivers (search/final_test.cpp, search/prepare_graph.cpp) compile against this
// header unchanged: `g++ -I gbnns_dim_red_b200/host -I include final_test.cpp -L... -lgbdr`.  Names,
// argument order, ownership (caller owns every std::vector, the callee borrows for the call) and
// printed/appended result lines follow the reference; the bodies hand the BATCH of queries to the
// C ABI in include/gbdr.h instead of looping over queries on the CPU:
//
//   performTest / performRealTests          search/search_function.h:128-210, 291-316  -> gbdr_search
//   performNetTest / performRealNetTests    search/search_function.h:319-408, 411-436  -> gbdr_search (q_low = NULL)
//   getOneSearchResults / getRealNearest    search/search_function.h:43-125            -> gbdr_search, batch of 1
//   GetLowQueryFromNet                      search/support_func.h:645-658              -> gbdr_project, batch of 1
//   hnswlikeGD                              search/support_func.h:521-575              -> gbdr_gd_prune
//   loadXvecs / loadEdges / writeEdges / readSearchParams / getVectorFromString
//                                           search/support_func.h:176-249, 578-621      (host file IO)
//
// use_second_graph == true (the KL "long link" experiments of naive_test.cpp, search_function.h:73-89)
// uploads the auxiliary graph with gbdr_index_set_aux_graph and searches with GBDR_SEARCH_SECOND_GRAPH.
//
// Deliberate differences, all on error paths or dead features:
//   * number_of_threads is accepted and ignored (the GPU runs the whole batch).
//   * a missing data file is an error (the reference silently reads zeros).
//   * there is no CPU fallback: without a B200 every call fails with the library's message.
#ifndef GBDR_HOST_SEARCH_FUNCTION_H_
#define GBDR_HOST_SEARCH_FUNCTION_H_

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "gbdr.h"

using namespace std;  // the reference exports its API at global scope with this in effect (search_function.h:4)

// ------------------------------------------------------------------------------------------------
// plumbing shared by the wrappers
// ------------------------------------------------------------------------------------------------
namespace gbdr_host {

[[noreturn]] inline void die(const string& what) {
    cout << "gbdr: " << what << endl;
    exit(1);
}
inline void check(int rc, const char* what) {
    if (rc != GBDR_OK) die(string(what) + ": " + gbdr_last_error());
}
inline int device() {
    const char* s = getenv("GBDR_DEVICE");
    return s && *s ? atoi(s) : 0;
}
// GBDR_DEVICES="0,1,2,3": the batched entry points (performTest / performNetTest) replicate the index on these GPUs and
// split every batch over them (gbdr_group_*), the GPU-side form of the reference's OpenMP team over the queries
// (search_function.h:147-152).  Unset or a single id: one GPU (GBDR_DEVICE).
inline vector<int> devices() {
    vector<int> out;
    const char* s = getenv("GBDR_DEVICES");
    if (s && *s) {
        string tok;
        istringstream ss(s);
        while (getline(ss, tok, ','))
            if (!tok.empty()) out.push_back(atoi(tok.c_str()));
    }
    if (out.empty()) out.push_back(device());
    return out;
}

struct FlatGraph {
    vector<uint64_t> offsets;
    vector<uint32_t> edges;
};
inline FlatGraph flatten(const vector<vector<uint32_t>>& g) {
    FlatGraph f;
    f.offsets.resize(g.size() + 1);
    f.offsets[0] = 0;
    for (size_t i = 0; i < g.size(); ++i) f.offsets[i + 1] = f.offsets[i] + g[i].size();
    f.edges.resize(f.offsets.back());
    for (size_t i = 0; i < g.size(); ++i)
        if (!g[i].empty()) memcpy(f.edges.data() + f.offsets[i], g[i].data(), g[i].size() * sizeof(uint32_t));
    return f;
}

// content hash of a whole buffer (four independent multiply-xor lanes over 8-byte words, ~10 GB/s): a vector that was
// refilled or edited in place, however sparsely, is uploaded again
inline uint64_t fingerprint(const void* p, size_t bytes) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    uint64_t h[4] = {0x9E3779B97F4A7C15ull ^ bytes, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
    size_t i = 0;
    for (; i + 32 <= bytes; i += 32) {
        uint64_t w[4];
        memcpy(w, b + i, 32);
        for (int j = 0; j < 4; ++j) {
            h[j] = (h[j] ^ w[j]) * 0x100000001B3ull;
            h[j] ^= h[j] >> 29;
        }
    }
    uint64_t t = 0;
    for (; i < bytes; ++i) t = (t << 8) ^ b[i] ^ (t >> 56);
    uint64_t r = t;
    for (int j = 0; j < 4; ++j) r = (r ^ h[j]) * 0xFF51AFD7ED558CCDull, r ^= r >> 33;
    return r;
}

// what the wrappers search: one index on one GPU, or a replicated group over GBDR_DEVICES
struct Resident {
    gbdr_index* ix = nullptr;
    gbdr_group* grp = nullptr;
    vector<gbdr_index*> views;  // single GPU: extra handles so that repetitions of a batch can be in flight together
    uint64_t last_use = 0;
    void destroy() {
        for (gbdr_index* v : views) gbdr_index_destroy(v);
        views.clear();
        if (ix) gbdr_index_destroy(ix);
        if (grp) gbdr_group_destroy(grp);
        ix = nullptr;
        grp = nullptr;
    }
};

// One resident index per distinct (base, low-dim base, graph, net) combination seen by the wrappers; at most four are
// kept, the least recently used one is dropped when a fifth arrives.
class IndexCache {
  public:
    struct Key {
        uint64_t base = 0, low = 0, graph = 0, net = 0, multi = 0;
        bool operator<(const Key& o) const {
            return tie(base, low, graph, net, multi) < tie(o.base, o.low, o.graph, o.net, o.multi);
        }
    };
    static IndexCache& instance() {
        static IndexCache c;
        return c;
    }
    static uint64_t graph_fingerprint(const vector<vector<uint32_t>>& graph) {
        uint64_t h = graph.size();
        for (size_t i = 0; i < graph.size(); ++i)
            h = (h * 1099511628211ull) ^ fingerprint(graph[i].data(), graph[i].size() * 4) ^ graph[i].size();
        return h;
    }
    // the auxiliary graph of use_second_graph searches, uploaded once per (index, graph, hops_bound, llf)
    void set_aux(Resident* r, const vector<vector<uint32_t>>& aux, uint32_t hops_bound, bool llf) {
        const uint64_t fp = graph_fingerprint(aux) * 31u + hops_bound * 2u + (llf ? 1u : 0u);
        auto it = aux_.find(r);
        if (it != aux_.end() && it->second == fp) return;
        FlatGraph f = flatten(aux);
        const int members = r->grp ? gbdr_group_size(r->grp) : 1;
        for (int i = 0; i < members; ++i) {
            gbdr_index* h = r->ix;
            if (r->grp) check(gbdr_group_member(r->grp, i, &h), "gbdr_group_member");
            check(gbdr_index_set_aux_graph(h, f.offsets.data(), f.edges.data(), aux.size(), hops_bound, llf ? 1 : 0),
                  "gbdr_index_set_aux_graph");
        }
        aux_[r] = fp;
    }
    // multi: use every GPU of GBDR_DEVICES (batched entry points); otherwise one index on GBDR_DEVICE
    Resident* get(const float* base, size_t n, size_t d, const float* low, size_t d_low,
                  const vector<vector<uint32_t>>* graph, const float* l1, const float* l2, const float* l3,
                  size_t net_d, size_t dh, size_t dh2, size_t net_dlow, bool multi = false) {
        const vector<int> devs = devices();
        multi = multi && devs.size() > 1;
        Key k;
        if (base) k.base = fingerprint(base, n * d * sizeof(float));
        if (low) k.low = fingerprint(low, n * d_low * sizeof(float));
        if (graph) k.graph = graph_fingerprint(*graph);
        if (l1) k.net = fingerprint(l1, dh * (net_d + 1) * 4) ^ fingerprint(l2, dh2 * (dh + 1) * 4) * 3u ^
                        fingerprint(l3, net_dlow * (dh2 + 1) * 4) * 5u;
        k.multi = multi ? devs.size() : 0;
        auto it = map_.find(k);
        if (it != map_.end()) {
            it->second.last_use = ++clock_;
            return &it->second;
        }
        if (map_.size() >= 4) {  // keep HBM bounded: drop the least recently used index
            auto victim = map_.begin();
            for (auto jt = map_.begin(); jt != map_.end(); ++jt)
                if (jt->second.last_use < victim->second.last_use) victim = jt;
            aux_.erase(&victim->second);
            victim->second.destroy();
            map_.erase(victim);
        }
        Resident r;
        FlatGraph f;
        if (graph) f = flatten(*graph);
        if (multi) {
            check(gbdr_group_create(devs.data(), (int)devs.size(), GBDR_GROUP_REPLICATED, &r.grp), "gbdr_group_create");
            if (base) check(gbdr_group_set_base(r.grp, base, n, (uint32_t)d), "gbdr_group_set_base");
            if (low) check(gbdr_group_set_low(r.grp, low, n, (uint32_t)d_low), "gbdr_group_set_low");
            if (graph) check(gbdr_group_set_graph(r.grp, f.offsets.data(), f.edges.data(), graph->size()), "gbdr_group_set_graph");
            if (l1)
                check(gbdr_group_set_net(r.grp, l1, l2, l3, (uint32_t)net_d, (uint32_t)dh, (uint32_t)dh2, (uint32_t)net_dlow),
                      "gbdr_group_set_net");
        } else {
            check(gbdr_index_create(device(), &r.ix), "gbdr_index_create");
            if (base) check(gbdr_index_set_base(r.ix, base, n, (uint32_t)d), "gbdr_index_set_base");
            if (low) check(gbdr_index_set_low(r.ix, low, n, (uint32_t)d_low), "gbdr_index_set_low");
            if (graph) check(gbdr_index_set_graph(r.ix, f.offsets.data(), f.edges.data(), graph->size()), "gbdr_index_set_graph");
            if (l1)
                check(gbdr_index_set_net(r.ix, l1, l2, l3, (uint32_t)net_d, (uint32_t)dh, (uint32_t)dh2, (uint32_t)net_dlow),
                      "gbdr_index_set_net");
        }
        r.last_use = ++clock_;
        return &(map_[k] = r);
    }
    // forget every resident copy (a caller that changed its vectors behind a const pointer may also just call this)
    void clear() {
        for (auto& kv : map_) kv.second.destroy();
        map_.clear();
        aux_.clear();
    }
    ~IndexCache() { clear(); }

  private:
    map<Key, Resident> map_;
    map<Resident*, uint64_t> aux_;
    uint64_t clock_ = 0;
};

}  // namespace gbdr_host

// ------------------------------------------------------------------------------------------------
// support_classes.h: stopwatch
// ------------------------------------------------------------------------------------------------
class StopW {
    chrono::steady_clock::time_point time_begin;

  public:
    StopW() : time_begin(chrono::steady_clock::now()) {}
    float getElapsedTimeMicro() {
        return (float)chrono::duration_cast<chrono::microseconds>(chrono::steady_clock::now() - time_begin).count();
    }
    void reset() { time_begin = chrono::steady_clock::now(); }
};

// ------------------------------------------------------------------------------------------------
// visited_list_pool.h: the epoch-stamped visited array and its pool, for callers that drive makeStep themselves.
// The batched GPU path never touches these: the beam kernel owns its visited set (shared-memory table + HBM spill).
// ------------------------------------------------------------------------------------------------
typedef uint16_t vl_type;

class VisitedList {
  public:
    vl_type curV;
    vl_type* mass;
    size_t numelements;

    explicit VisitedList(size_t count) : curV((vl_type)-1), mass(new vl_type[count]), numelements(count) {}
    VisitedList(const VisitedList&) = delete;
    VisitedList& operator=(const VisitedList&) = delete;
    ~VisitedList() { delete[] mass; }
    // next epoch; the stamps are cleared when the 16-bit epoch wraps (visited_list_pool.h:21-28)
    void reset() {
        if (++curV == 0) {
            fill(mass, mass + numelements, (vl_type)0);
            curV = 1;
        }
    }
};

class VisitedListPool {
    vector<VisitedList*> idle_;
    mutex guard_;
    size_t numelements_;

  public:
    VisitedListPool(size_t initmaxpools, size_t numelements) : numelements_(numelements) {
        for (size_t i = 0; i < initmaxpools; ++i) idle_.push_back(new VisitedList(numelements));
    }
    VisitedListPool(const VisitedListPool&) = delete;
    VisitedListPool& operator=(const VisitedListPool&) = delete;
    ~VisitedListPool() {
        for (VisitedList* v : idle_) delete v;
    }
    VisitedList* getFreeVisitedList() {
        VisitedList* v = nullptr;
        {
            lock_guard<mutex> lock(guard_);
            if (!idle_.empty()) {
                v = idle_.back();
                idle_.pop_back();
            }
        }
        if (!v) v = new VisitedList(numelements_);
        v->reset();
        return v;
    }
    void releaseVisitedList(VisitedList* v) {
        lock_guard<mutex> lock(guard_);
        idle_.push_back(v);
    }
};

// ------------------------------------------------------------------------------------------------
// support_func.h: metrics (host-side scoring helpers; the search itself never calls these)
// ------------------------------------------------------------------------------------------------
struct Net {
    vector<float> layerFirst;
    vector<float> layerSecond;
    vector<float> layerFinal;
};

class Metric {
  public:
    virtual float Dist(const float* x, const float* y, size_t d) = 0;
    virtual ~Metric() {}
};

// squared L2 over floor(d/4)*4 dims, four lane-strided sums (support_func.h:107-128)
class L2Metric : public Metric {
  public:
    float Dist(const float* x, const float* y, size_t d) override {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (size_t c = 0; c + 4 <= d; c += 4)
            for (int j = 0; j < 4; ++j) {
                const float e = x[c + j] - y[c + j];
                s[j] = s[j] + e * e;
            }
        return ((s[0] + s[1]) + s[2]) + s[3];
    }
};

// NEGATED dot product (support_func.h:131-163)
class Angular : public Metric {
  public:
    float Dist(const float* x, const float* y, size_t d) override {
        float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        size_t i = 0;
        for (; i + 8 <= d; i += 8)
            for (int j = 0; j < 8; ++j) m[j] = m[j] + x[i + j] * y[i + j];
        float s[4];
        for (int j = 0; j < 4; ++j) s[j] = m[4 + j] + m[j];
        if (i + 4 <= d) {
            for (int j = 0; j < 4; ++j) s[j] = s[j] + x[i + j] * y[i + j];
            i += 4;
        }
        for (int j = 0; i < d; ++i, ++j) s[j] = s[j] + x[i] * y[i];
        return -((s[0] + s[1]) + (s[2] + s[3]));
    }
};

inline int findGraphAverageDegree(vector<vector<uint32_t>>& graph) {
    double total = 0;
    for (const auto& row : graph) total += row.size();
    return graph.empty() ? 0 : (int)(total / graph.size());
}

// ------------------------------------------------------------------------------------------------
// support_func.h: file formats (byte-identical with the reference and dim_red/data.py)
// ------------------------------------------------------------------------------------------------
template <typename T>
void readXvec(ifstream& in, T* data, const size_t d, const size_t n = 1) {
    for (size_t i = 0; i < n; i++) {
        uint32_t dim = 0;
        in.read((char*)&dim, sizeof(uint32_t));
        if (!in || dim != d) {
            cout << "file error\n";
            cout << "dim " << dim << ", d " << d << endl;
            exit(1);
        }
        in.read((char*)(data + i * d), d * sizeof(T));
    }
}

template <typename T>
void writeXvec(ofstream& out, T* data, const size_t d, const size_t n = 1) {
    const uint32_t dim = (uint32_t)d;
    for (size_t i = 0; i < n; i++) {
        out.write((char*)&dim, sizeof(uint32_t));
        out.write((char*)(data + i * d), d * sizeof(T));
    }
}

template <typename T>
vector<T> loadXvecs(string dataPath, const size_t d, const size_t n = 1) {
    vector<T> data(n * d);
    ifstream in(dataPath.c_str(), ios::binary);
    if (!in) gbdr_host::die("cannot open " + dataPath);
    readXvec<T>(in, data.data(), d, n);
    return data;
}

inline void writeEdges(string location, const vector<vector<uint32_t>>& edges) {
    cout << "Saving edges to " << location << endl;
    ofstream out(location.c_str(), ios::binary);
    if (!out) gbdr_host::die("cannot write " + location);
    for (const auto& row : edges) {
        const uint32_t size = (uint32_t)row.size();
        out.write((const char*)&size, sizeof(uint32_t));
        out.write((const char*)row.data(), sizeof(uint32_t) * size);
    }
}

inline vector<vector<uint32_t>> loadEdges(string location, uint32_t n, string edges_name) {
    vector<vector<uint32_t>> edges(n);
    ifstream in(location.c_str(), ios::binary);
    if (!in) gbdr_host::die("cannot open " + location);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t size = 0;
        in.read((char*)&size, sizeof(uint32_t));
        if (!in) gbdr_host::die("edge file " + location + " ends before vertex " + to_string(i));
        edges[i].resize(size);
        in.read((char*)edges[i].data(), sizeof(uint32_t) * size);
    }
    cout << edges_name + " " << findGraphAverageDegree(edges) << endl;
    return edges;
}

// "<dataset> <key> <value>" lines with exactly three space-separated tokens (support_func.h:578-610)
inline vector<string> splitString(const string& str, char sep) {
    vector<string> out;
    string tok;
    istringstream ss(str);
    while (getline(ss, tok, sep)) out.push_back(tok);
    return out;
}

inline map<string, string> readSearchParams(string fileName, string databaseName) {
    map<string, string> params;
    ifstream file(fileName);
    string line;
    while (getline(file, line)) {
        vector<string> tok = splitString(line, ' ');
        if (tok.size() == 3 && tok[0] == databaseName) params[tok[1]] = tok[2];
    }
    return params;
}

inline vector<int> getVectorFromString(string str) {
    vector<int> out;
    for (const string& t : splitString(str, ',')) out.push_back(atoi(t.c_str()));
    return out;
}

// ------------------------------------------------------------------------------------------------
// support_func.h: graph pruning and the projection net
// ------------------------------------------------------------------------------------------------
inline vector<vector<uint32_t>> hnswlikeGD(vector<vector<uint32_t>>& graph, const float* ds, int M, size_t N, size_t d,
                                           Metric* /*metric*/, bool reverse, bool need_const_degree) {
    gbdr_host::FlatGraph f = gbdr_host::flatten(graph);
    vector<uint64_t> off(N + 1);
    vector<uint32_t> edges(N * 2 * (size_t)M);
    double secs = 0;
    gbdr_host::check(gbdr_gd_prune(gbdr_host::device(), f.offsets.data(), f.edges.data(), ds, N, (uint32_t)d, (uint32_t)M,
                                   reverse, need_const_degree, off.data(), edges.data(), &secs),
                     "gbdr_gd_prune");
    vector<vector<uint32_t>> out(N);
    for (size_t i = 0; i < N; ++i) out[i].assign(edges.begin() + off[i], edges.begin() + off[i + 1]);
    return out;
}

// every list re-ranked by distance to its own vertex and cut to the knn_size nearest (support_func.h:309-340); the
// "fixed" constant-degree graphs of the README are made this way from the kNN-1k file
inline vector<vector<uint32_t>> cutKNNbyK(vector<vector<uint32_t>>& knn, const float* ds, int knn_size, int N, int d,
                                          Metric* /*metric*/) {
    for (int i = 0; i < N; ++i)
        if ((size_t)knn_size > knn[i].size()) {  // the reference's warning, printed once (:322-327)
            cout << "Size knn less than you want" << endl;
            cout << knn[i].size() << endl;
            break;
        }
    gbdr_host::FlatGraph f = gbdr_host::flatten(knn);
    vector<uint64_t> off((size_t)N + 1);
    vector<uint32_t> edges((size_t)N * (size_t)knn_size);
    gbdr_host::check(gbdr_knn_cut(gbdr_host::device(), f.offsets.data(), f.edges.data(), ds, (uint64_t)N, (uint32_t)d,
                                  (uint32_t)knn_size, off.data(), edges.data(), nullptr),
                     "gbdr_knn_cut");
    vector<vector<uint32_t>> out(N);
    for (int i = 0; i < N; ++i) out[i].assign(edges.begin() + off[i], edges.begin() + off[i + 1]);
    return out;
}

inline void GetLowQueryFromNet(const Net* net, const float* query, vector<float>& ans, const float* /*zeros*/, size_t d,
                               size_t d_hidden, size_t d_hidden_2, size_t d_low, Metric* /*ang*/, Metric* /*l2*/) {
    gbdr_host::Resident* r = gbdr_host::IndexCache::instance().get(nullptr, 0, 0, nullptr, 0, nullptr,
                                                                   net->layerFirst.data(), net->layerSecond.data(),
                                                                   net->layerFinal.data(), d, d_hidden, d_hidden_2, d_low);
    ans.resize(d_low);
    gbdr_host::check(gbdr_project(r->ix, query, 1, ans.data()), "gbdr_project");
}

// ------------------------------------------------------------------------------------------------
// search_function.h: per-query entry points (batch of one)
// ------------------------------------------------------------------------------------------------
struct TripleResult {
    priority_queue<pair<float, int>> topk;
    int hops;
    int dist_calc;
    int degree;
};

// One expansion step on the HOST, for callers that walk a graph themselves (search_function.h:15-40): every not yet
// visited neighbour is stamped, measured, and enters both heaps if it beats the worst of a full result heap.  The batched
// entry points do not come through here — their whole walk is the beam kernel.
inline void makeStep(vector<uint32_t>& graph_level, const float* query, const float* db,
                     priority_queue<pair<float, int>>& topResults, priority_queue<pair<float, int>>& candidateSet,
                     Metric* metric, uint32_t d, int& query_dist_calc, bool& found, int& ef, int& /*k*/,
                     VisitedList* vl) {
    vl_type* const stamp = vl->mass;
    const vl_type epoch = vl->curV;
    for (uint32_t nb : graph_level) {
        if (stamp[nb] == epoch) continue;
        stamp[nb] = epoch;
        const float dist = metric->Dist(query, db + (size_t)nb * d, d);
        ++query_dist_calc;
        if (topResults.size() < (size_t)ef || topResults.top().first > dist) {
            candidateSet.emplace(-dist, (int)nb);
            topResults.emplace(dist, (int)nb);
            found = true;
            if (topResults.size() > (size_t)ef) topResults.pop();
        }
    }
}

inline TripleResult getOneSearchResults(const float* query, const float* db, uint32_t N, uint32_t d,
                                        vector<vector<uint32_t>>& main_graph, vector<vector<uint32_t>>& auxiliary_graph,
                                        int ef, int k, vector<uint32_t>& inter_points, Metric* /*metric*/,
                                        VisitedListPool* /*visitedlistpool*/, bool use_second_graph, bool llf,
                                        uint32_t hops_bound) {
    if (inter_points.size() != 1) gbdr_host::die("exactly one entry point per query is supported");
    gbdr_host::Resident* r = gbdr_host::IndexCache::instance().get(nullptr, N, 0, db, d, &main_graph, nullptr, nullptr,
                                                                   nullptr, 0, 0, 0, 0);
    if (use_second_graph) gbdr_host::IndexCache::instance().set_aux(r, auxiliary_graph, hops_bound, llf);
    gbdr_index* h = r->ix;
    vector<uint32_t> ids(k);
    vector<float> dists(k);
    TripleResult tr;
    tr.degree = 0;
    gbdr_host::check(gbdr_search(h, nullptr, query, 1, (uint32_t)ef, (uint32_t)k,
                                 use_second_graph ? GBDR_SEARCH_SECOND_GRAPH : 0u, inter_points.data(), ids.data(),
                                 dists.data(), &tr.hops, &tr.dist_calc, nullptr),
                     "gbdr_search");
    for (int j = 0; j < k; ++j)
        if (ids[j] != GBDR_PAD_ID) tr.topk.push(make_pair(dists[j], (int)ids[j]));
    return tr;
}

// exact re-rank of the low-dimensional survivors; walks the heap from its worst element and keeps the
// strictly smaller distance (search_function.h:105-125).  A handful of rows: done on the host here, the
// batched path (performTest) runs it on the GPU inside gbdr_search.
inline int getRealNearest(const float* point_q, int /*k*/, int d, int /*d_low*/, priority_queue<pair<float, int>>& topk,
                          vector<float>& ds, Metric* metric) {
    int best = topk.top().second;
    float best_dist = metric->Dist(ds.data() + (size_t)d * best, point_q, d);
    topk.pop();
    while (!topk.empty()) {
        const int id = topk.top().second;
        const float dist = metric->Dist(ds.data() + (size_t)d * id, point_q, d);
        if (dist < best_dist) {
            best_dist = dist;
            best = id;
        }
        topk.pop();
    }
    return best;
}

// ------------------------------------------------------------------------------------------------
// search_function.h: the batched harness
// ------------------------------------------------------------------------------------------------
namespace gbdr_host {

struct BatchStats {
    long long hops = 0, dist_calc = 0;
    double acc = 0, work_time_us = 0;
    int num_exp = 0;
};

// page-locks a caller-owned vector for the duration of a batch run so that uploads are real asynchronous DMA
struct PinScope {
    vector<void*> regs;
    void pin(const void* p, size_t bytes) {
        if (p && bytes && gbdr_host_register(const_cast<void*>(p), bytes) == GBDR_OK) regs.push_back(const_cast<void*>(p));
    }
    ~PinScope() {
        for (void* p : regs) gbdr_host_unregister(p);
    }
};

template <typename T>
struct PinnedArray {  // page-locked result buffer (D2H copies overlap the next repetition's kernels)
    T* p = nullptr;
    explicit PinnedArray(size_t n) { check(gbdr_host_alloc_pinned(std::max<size_t>(n, 1) * sizeof(T), (void**)&p), "gbdr_host_alloc_pinned"); }
    ~PinnedArray() { gbdr_host_free_pinned(p); }
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    T& operator[](size_t i) { return p[i]; }
};

// one ef point: `number_exper` timed repetitions of the whole query batch, scored like the reference
// (search_function.h:147-203).  The repetitions are independent, so on one GPU up to three of them are in flight at
// a time (the index and two views of it, gbdr_search_submit / gbdr_search_wait): uploads, kernels and downloads of
// neighbouring repetitions overlap the way bench.py's e2e leg does it.  With a group every repetition is one
// gbdr_group_search over all GPUs of GBDR_DEVICES.
inline BatchStats run_point(Resident* r, vector<float>& ds, vector<float>& queries, const float* queries_low,
                            vector<uint32_t>& truth, int n, int d, int d_low, int n_q, int n_tr, int ef, int k,
                            Metric* metric, const vector<vector<uint32_t>>& inter_points, int dist_calc_boost,
                            int recheck_size, int number_exper, bool net_mode, bool second_graph = false) {
    (void)n;
    PinnedArray<uint32_t> entry(n_q);
    for (int i = 0; i < n_q; ++i) {
        if (inter_points[i].size() != 1) die("exactly one entry point per query is supported");
        entry[i] = inter_points[i][0];
    }
    BatchStats st;
    st.dist_calc = (long long)dist_calc_boost * n_q;
    const bool low_dim = d != d_low;
    const bool rerank = low_dim && recheck_size > 0;
    const uint32_t beam = rerank ? (uint32_t)recheck_size : (uint32_t)ef;
    const uint32_t kk = rerank ? 1u : (uint32_t)k;
    const uint32_t flags = (rerank ? GBDR_SEARCH_RERANK : (low_dim ? 0u : GBDR_SEARCH_PLAIN)) |
                           (second_graph ? GBDR_SEARCH_SECOND_GRAPH : 0u);
    const float* q_low = low_dim && !net_mode ? queries_low : nullptr;

    vector<gbdr_index*> lanes;  // handles that can each hold one repetition in flight
    if (!r->grp) {
        lanes.push_back(r->ix);
        const size_t want = (size_t)std::min(std::max(number_exper, 1), 3);
        while (1 + r->views.size() < want) {
            gbdr_index* v = nullptr;
            check(gbdr_index_create_view(r->ix, &v), "gbdr_index_create_view");
            r->views.push_back(v);
        }
        for (size_t j = 0; j + 1 < want; ++j) lanes.push_back(r->views[j]);
    }
    const size_t L = r->grp ? 1 : lanes.size();
    vector<unique_ptr<PinnedArray<uint32_t>>> ids(L);
    vector<unique_ptr<PinnedArray<int32_t>>> hops(L), dcs(L);
    for (size_t j = 0; j < L; ++j) {
        ids[j].reset(new PinnedArray<uint32_t>((size_t)n_q * kk));
        hops[j].reset(new PinnedArray<int32_t>(n_q));
        dcs[j].reset(new PinnedArray<int32_t>(n_q));
    }
    PinScope pins;
    pins.pin(queries.data(), queries.size() * sizeof(float));
    if (q_low) pins.pin(q_low, (size_t)n_q * d_low * sizeof(float));

    // results of every repetition, scored after the clock has stopped (the reference stops its clock before scoring,
    // :187-189; here a repetition's scoring would otherwise overlap the next one's kernels)
    vector<uint32_t> all_ids((size_t)std::max(number_exper, 0) * n_q * kk);
    vector<int32_t> all_hops((size_t)std::max(number_exper, 0) * n_q), all_dcs((size_t)std::max(number_exper, 0) * n_q);
    auto keep = [&](size_t j, int v) {  // lane buffers -> repetition v's slot (120 KB: microseconds)
        memcpy(all_ids.data() + (size_t)v * n_q * kk, ids[j]->p, (size_t)n_q * kk * sizeof(uint32_t));
        memcpy(all_hops.data() + (size_t)v * n_q, hops[j]->p, (size_t)n_q * sizeof(int32_t));
        memcpy(all_dcs.data() + (size_t)v * n_q, dcs[j]->p, (size_t)n_q * sizeof(int32_t));
    };
    auto score = [&](int v) {
        st.num_exp += 1;
        for (int i = 0; i < n_q; ++i) {
            // `while (topk.size() > k) pop; ans = topk.top().second`: the worst of the k best (:168-171)
            const uint32_t* row = all_ids.data() + ((size_t)v * n_q + i) * kk;
            uint32_t a = row[0];
            for (uint32_t t = 1; t < kk; ++t)
                if (row[t] != GBDR_PAD_ID) a = row[t];
            st.hops += all_hops[(size_t)v * n_q + i];
            st.dist_calc += all_dcs[(size_t)v * n_q + i];
            st.acc += a == truth[(size_t)i * n_tr];
            if (n_tr > 1) {  // duplicated ground-truth vectors in SIFT (:193-202)
                const float* x = ds.data() + (size_t)d * truth[(size_t)i * n_tr];
                const float* y = ds.data() + (size_t)d * truth[(size_t)i * n_tr + 1];
                if (metric->Dist(x, y, d) == 0 && truth[(size_t)i * n_tr] != truth[(size_t)i * n_tr + 1])
                    st.acc += a == truth[(size_t)i * n_tr + 1];
            }
        }
    };

    // workspaces of every lane are sized by its first call at this (n_q, ef): do that call before the clock starts (index
    // residency is set-up, like the upload of the base vectors), once per lane and shape
    static map<pair<void*, uint64_t>, bool> warmed;
    const uint64_t shape_key = ((uint64_t)n_q << 32) ^ ((uint64_t)beam << 8) ^ flags;
    if (r->grp) {
        if (!warmed[make_pair((void*)r->grp, shape_key)]) {
            check(gbdr_group_search(r->grp, queries.data(), q_low, (uint32_t)n_q, beam, kk, flags, entry.p, ids[0]->p, nullptr,
                                    hops[0]->p, dcs[0]->p, nullptr),
                  "gbdr_group_search");
            warmed[make_pair((void*)r->grp, shape_key)] = true;
        }
    } else {
        for (size_t j = 0; j < L; ++j)
            if (!warmed[make_pair((void*)lanes[j], shape_key)]) {
                check(gbdr_search(lanes[j], queries.data(), q_low, (uint32_t)n_q, beam, kk, flags, entry.p, ids[j]->p, nullptr,
                                  hops[j]->p, dcs[j]->p, nullptr),
                      "gbdr_search");
                warmed[make_pair((void*)lanes[j], shape_key)] = true;
            }
    }
    StopW stopw;
    if (r->grp) {
        for (int v = 0; v < number_exper; ++v) {
            check(gbdr_group_search(r->grp, queries.data(), q_low, (uint32_t)n_q, beam, kk, flags, entry.p, ids[0]->p, nullptr,
                                    hops[0]->p, dcs[0]->p, nullptr),
                  "gbdr_group_search");
            keep(0, v);
        }
    } else {
        int submitted = 0, waited = 0;
        while (waited < number_exper) {
            for (; submitted < number_exper && submitted - waited < (int)L; ++submitted) {
                const size_t j = (size_t)submitted % L;
                check(gbdr_search_submit(lanes[j], queries.data(), q_low, (uint32_t)n_q, beam, kk, flags, entry.p, ids[j]->p,
                                         nullptr, hops[j]->p, dcs[j]->p),
                      "gbdr_search_submit");
            }
            const size_t j = (size_t)waited % L;
            check(gbdr_search_wait(lanes[j], nullptr), "gbdr_search_wait");
            keep(j, waited);
            ++waited;
        }
    }
    st.work_time_us = stopw.getElapsedTimeMicro();
    for (int v = 0; v < number_exper; ++v) score(v);
    return st;
}

inline void report(const BatchStats& st, int n_q, const string& graph_name, const char* output_txt) {
    const double denom = (double)st.num_exp * n_q;
    ostringstream line;  // same tokens as search_function.h:206-209 (hops and dist_calc are integer divisions)
    line << "graph_type " << graph_name << " acc " << (float)(st.acc / denom) << " hops " << st.hops / (long long)denom
         << " dist_calc " << st.dist_calc / (long long)denom << " work_time " << (float)(st.work_time_us / (denom * 1e6));
    cout << line.str() << endl;
    ofstream out(output_txt, ios_base::app);
    out << line.str() << endl;
}

inline vector<vector<uint32_t>> make_entry_points(int n, int n_q, mt19937& random_gen, const string& graph_name) {
    // graphs whose name starts with "hnsw" enter at vertex 0, every other graph at a uniform random
    // vertex per query (search_function.h:297-307)
    vector<vector<uint32_t>> inter_points(n_q);
    const int mult = graph_name.substr(0, 4) == "hnsw" ? 0 : 1;
    uniform_int_distribution<int> uniform_distr(0, n - 1);
    for (int j = 0; j < n_q; ++j) inter_points[j].push_back((uint32_t)(uniform_distr(random_gen) * mult));
    return inter_points;
}

}  // namespace gbdr_host

inline void performTest(vector<vector<uint32_t>>& knn_graph, vector<vector<uint32_t>>& kl_graph, vector<float>& ds,
                        vector<float>& queries, vector<float>& ds_low, vector<float>& queries_low, vector<uint32_t>& truth,
                        int n, int d, int d_low, int n_q, int n_tr, int ef, int k, string graph_name, Metric* metric,
                        const char* output_txt, vector<vector<uint32_t>> inter_points, bool use_second_graph, bool llf,
                        uint32_t hops_bound, int dist_calc_boost, int recheck_size, int number_exper,
                        int /*number_of_threads*/) {
    const bool low_dim = d != d_low;
    gbdr_host::Resident* h = gbdr_host::IndexCache::instance().get(ds.data(), n, d, low_dim ? ds_low.data() : nullptr,
                                                                   d_low, &knn_graph, nullptr, nullptr, nullptr, 0, 0, 0, 0,
                                                                   true);
    if (use_second_graph) gbdr_host::IndexCache::instance().set_aux(h, kl_graph, hops_bound, llf);
    gbdr_host::BatchStats st =
        gbdr_host::run_point(h, ds, queries, queries_low.data(), truth, n, d, d_low, n_q, n_tr, ef, k, metric, inter_points,
                             dist_calc_boost, recheck_size, number_exper, false, use_second_graph);
    gbdr_host::report(st, n_q, graph_name, output_txt);
}

inline void performRealTests(int n, int d, int d_low, int n_q, int n_tr, vector<int> efs, mt19937 random_gen,
                             vector<vector<uint32_t>>& main_graph, vector<vector<uint32_t>>& kl, vector<float>& db,
                             vector<float>& queries, vector<float>& db_low, vector<float>& queries_low,
                             vector<uint32_t>& truth, const char* output_txt, Metric* metric, string graph_name,
                             bool use_second_graph, bool llf, int number_exper, int number_of_threads) {
    vector<vector<uint32_t>> inter_points = gbdr_host::make_entry_points(n, n_q, random_gen, graph_name);
    const uint32_t hops_bound = 50;
    for (size_t i = 0; i < efs.size(); ++i)
        performTest(main_graph, kl, db, queries, db_low, queries_low, truth, n, d, d_low, n_q, n_tr, efs[i], 1, graph_name,
                    metric, output_txt, inter_points, use_second_graph, llf, hops_bound, 0, efs[i], number_exper,
                    number_of_threads);
}

inline void performNetTest(vector<vector<uint32_t>>& knn_graph, vector<vector<uint32_t>>& kl_graph, vector<float>& ds,
                           vector<float>& queries, vector<float>& ds_low, const Net* net, size_t d_hidden,
                           vector<uint32_t>& truth, int n, int d, int d_low, int n_q, int n_tr, int ef, int k,
                           string graph_name, Metric* metric, const char* output_txt,
                           vector<vector<uint32_t>> inter_points, bool use_second_graph, bool llf,
                           uint32_t hops_bound, int dist_calc_boost, int recheck_size, int number_exper,
                           int /*number_of_threads*/) {
    const bool low_dim = d != d_low;
    gbdr_host::Resident* h = gbdr_host::IndexCache::instance().get(
        ds.data(), n, d, low_dim ? ds_low.data() : nullptr, d_low, &knn_graph, low_dim ? net->layerFirst.data() : nullptr,
        net->layerSecond.data(), net->layerFinal.data(), d, d_hidden, d_hidden, d_low, true);
    if (use_second_graph) gbdr_host::IndexCache::instance().set_aux(h, kl_graph, hops_bound, llf);
    gbdr_host::BatchStats st =
        gbdr_host::run_point(h, ds, queries, nullptr, truth, n, d, d_low, n_q, n_tr, ef, k, metric, inter_points,
                             dist_calc_boost, recheck_size, number_exper, true, use_second_graph);
    gbdr_host::report(st, n_q, graph_name, output_txt);
}

inline void performRealNetTests(int n, int d, int d_low, int n_q, int n_tr, vector<int> efs, mt19937 random_gen,
                                vector<vector<uint32_t>>& main_graph, vector<vector<uint32_t>>& kl, vector<float>& db,
                                vector<float>& queries, vector<float>& db_low, const Net* net, size_t d_hidden,
                                vector<uint32_t>& truth, const char* output_txt, Metric* metric, string graph_name,
                                bool use_second_graph, bool llf, int number_exper, int number_of_threads) {
    vector<vector<uint32_t>> inter_points = gbdr_host::make_entry_points(n, n_q, random_gen, graph_name);
    const uint32_t hops_bound = 50;
    for (size_t i = 0; i < efs.size(); ++i)
        performNetTest(main_graph, kl, db, queries, db_low, net, d_hidden, truth, n, d, d_low, n_q, n_tr, efs[i], 1,
                       graph_name, metric, output_txt, inter_points, use_second_graph, llf, hops_bound, 0, efs[i],
                       number_exper, number_of_threads);
}

#endif  // GBDR_HOST_SEARCH_FUNCTION_H_
