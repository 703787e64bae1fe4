// Test driver for the drop-in header: the two use_second_graph sweeps of the reference's
// naive_test.cpp:102-105 (performRealTests with a long-link graph, llf = true) on files given by argv.
//   second_graph_driver <base.fvecs> <query.fvecs> <truth.ivecs> <base_low.fvecs> <query_low.fvecs>
//                       <main.edges> <aux.edges> <out.txt> n d d_low n_q n_tr
#include "search_function.h"

int main(int argc, char** argv) {
    if (argc != 14) {
        cout << " Need to specify parameters" << endl;
        return 1;
    }
    const size_t n = atoi(argv[9]), d = atoi(argv[10]), d_low = atoi(argv[11]), n_q = atoi(argv[12]), n_tr = atoi(argv[13]);
    L2Metric l2 = L2Metric();
    std::mt19937 random_gen(1);
    vector<float> db = loadXvecs<float>(argv[1], d, n);
    vector<float> queries = loadXvecs<float>(argv[2], d, n_q);
    vector<uint32_t> truth = loadXvecs<uint32_t>(argv[3], n_tr, n_q);
    vector<float> db_low = loadXvecs<float>(argv[4], d_low, n);
    vector<float> queries_low = loadXvecs<float>(argv[5], d_low, n_q);
    vector<vector<uint32_t>> graph = loadEdges(argv[6], n, "main");
    vector<vector<uint32_t>> kl = loadEdges(argv[7], n, "kl");
    vector<int> efs = {6, 30};
    remove(argv[8]);
    // graph names starting with "hnsw" enter at vertex 0 (search_function.h:297-307): deterministic
    performRealTests(n, d, d_low, n_q, n_tr, efs, random_gen, graph, kl, db, queries, db_low, queries_low, truth, argv[8],
                     &l2, "hnsw_lk_low", true, true, 1, 1);
    performRealTests(n, d, d_low, n_q, n_tr, efs, random_gen, graph, kl, db, queries, db_low, queries_low, truth, argv[8],
                     &l2, "hnsw_lk_low_nollf", true, false, 1, 1);
    performRealTests(n, d, d_low, n_q, n_tr, efs, random_gen, graph, graph, db, queries, db_low, queries_low, truth, argv[8],
                     &l2, "hnsw_low", false, false, 1, 1);
    return 0;
}
