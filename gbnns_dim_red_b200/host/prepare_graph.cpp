// prepare_graph — drop-in for the reference's search/prepare_graph.cpp: `./prepare_graph <dataset> <latName>`.
//
// Loads the transformed base `<ds>_base_<lat>.fvecs` and the wide kNN lists `<ds>_knn_1k_<lat>.ivecs`,
// prunes them with hnswlikeGD(M = 30, reverse edges, no constant degree) on the GPU and writes
// `<ds>_gd_knn_<lat>.ivecs` (prepare_graph.cpp:64-74).  In the reference the kNN-1k file comes from the
// Python side (dim_red/support_func.py:374-384); when it is missing, or with GBDR_BUILD_KNN=1, it is
// built here by the GPU brute-force kNN (gbdr_knn, self at rank 0, k = GBDR_KNN_K, default 1000) and
// written in the same ivecs format first.
// Paths: GBDR_PARAMS, GBDR_DATA_ROOT, GBDR_MODELS_ROOT as in final_test; GBDR_GD_M overrides M.
#include "search_function.h"

static string env_or(const char* name, const string& dflt) {
    const char* s = getenv(name);
    return s && *s ? string(s) : dflt;
}

int main(int argc, char** argv) {
    string datasetName, fileLatName;
    if (argc == 3) {
        datasetName = argv[1];
        fileLatName = argv[2];
    } else {
        cout << " Need to specify parameters" << endl;
        return 1;
    }
    L2Metric l2;
    const string params_path = env_or("GBDR_PARAMS", "/home/shekhale/gbnns_dim_red/search/parameters_of_databases.txt");
    map<string, string> params_map = readSearchParams(params_path, datasetName);
    const size_t n = atoi(params_map["n"].c_str());
    const size_t n_q = atoi(params_map["n_q"].c_str());
    const size_t n_tr = atoi(params_map["n_tr"].c_str());
    const size_t d = atoi(params_map["d"].c_str());
    const size_t d_low = atoi(params_map["d_low"].c_str());
    cout << n << " " << n_q << " " << n_tr << " " << d << " " << d_low << endl;
    if (!n || !d_low) {
        cout << "dataset " << datasetName << " not described in " << params_path << endl;
        return 1;
    }
    const string pathData = env_or("GBDR_DATA_ROOT", "/mnt/data/shekhale/data") + "/" + datasetName + "/" + datasetName;
    const string pathModels =
        env_or("GBDR_MODELS_ROOT", "/mnt/data/shekhale/models/nns_graphs") + "/" + datasetName + "/" + datasetName;

    vector<float> db_low = loadXvecs<float>(pathData + "_base_" + fileLatName + ".fvecs", d_low, n);

    const string knnPath = pathModels + "_knn_1k_" + fileLatName + ".ivecs";
    const bool have_knn = (bool)ifstream(knnPath.c_str(), ios::binary);
    const int M = atoi(env_or("GBDR_GD_M", "30").c_str());
    if (!have_knn || env_or("GBDR_BUILD_KNN", "0") == "1") {
        // No kNN file yet (the reference gets it from its Python side, dim_red/triplet.py:266-268): build both graphs in
        // one chain that never leaves HBM — kNN self-join, hnswlikeGD on the lists where they lie — on one GPU, or
        // row-block sharded over the GPUs of GBDR_DEVICES.  The kNN lists are written out as the `_knn_1k_` file.
        const size_t k = min<size_t>(n, atoi(env_or("GBDR_KNN_K", "1000").c_str()));
        vector<uint32_t> ids(n * k), edges(n * 2 * (size_t)M);
        vector<uint64_t> off(n + 1);
        double t[4] = {0, 0, 0, 0};
        const vector<int> devs = gbdr_host::devices();
        if (devs.size() > 1) {
            gbdr_group* g = nullptr;
            gbdr_host::check(gbdr_group_create(devs.data(), (int)devs.size(), GBDR_GROUP_REPLICATED, &g), "gbdr_group_create");
            const int rc = gbdr_group_build_graph(g, db_low.data(), n, (uint32_t)d_low, (uint32_t)k, (uint32_t)M, 1, 0, off.data(),
                                                  edges.data(), ids.data(), t);
            gbdr_group_destroy(g);
            gbdr_host::check(rc, "gbdr_group_build_graph");
        } else {
            gbdr_host::check(gbdr_build_graph(gbdr_host::device(), db_low.data(), n, (uint32_t)d_low, (uint32_t)k, (uint32_t)M, 1, 0,
                                              off.data(), edges.data(), ids.data(), t),
                             "gbdr_build_graph");
        }
        cout << "knn_" << k << " built in " << t[1] << " s, GD graph in " << t[2] + t[3] << " s (" << devs.size() << " GPU)" << endl;
        {
            ofstream out(knnPath.c_str(), ios::binary);
            if (!out) gbdr_host::die("cannot write " + knnPath);
            writeXvec<uint32_t>(out, ids.data(), k, n);
        }
        cout << "knn_low " << k << endl;
        vector<vector<uint32_t>> gd(n);
        for (size_t i = 0; i < n; ++i) gd[i].assign(edges.begin() + off[i], edges.begin() + off[i + 1]);
        cout << "GD_knn " << findGraphAverageDegree(gd) << endl;
        writeEdges(pathModels + "_gd_knn_" + fileLatName + ".ivecs", gd);
        return 0;
    }
    vector<vector<uint32_t>> knn_low = loadEdges(knnPath, n, "knn_low");

    vector<vector<uint32_t>> gd_knn_low = hnswlikeGD(knn_low, db_low.data(), M, n, d_low, &l2, true, false);
    cout << "GD_knn " << findGraphAverageDegree(gd_knn_low) << endl;
    writeEdges(pathModels + "_gd_knn_" + fileLatName + ".ivecs", gd_knn_low);
    return 0;
}
