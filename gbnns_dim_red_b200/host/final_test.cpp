// final_test — drop-in for the reference's search/final_test.cpp: `./final_test <dataset>`.
//
// Reads the same parameter file and the same data / graph / net files, runs the original-dimension
// baseline sweep (performRealTests) and the dimensionality-reduction sweep with on-the-fly query
// projection (performRealNetTests), prints and appends the same result lines.  The reference
// hard-codes its author's directories (final_test.cpp:26,44-45,80); here they are the defaults of
// environment variables:
//   GBDR_PARAMS        parameter file            (/home/shekhale/gbnns_dim_red/search/parameters_of_databases.txt)
//   GBDR_DATA_ROOT     <root>/<ds>/<ds>_*.fvecs  (/mnt/data/shekhale/data)
//   GBDR_MODELS_ROOT   <root>/<ds>/...           (/mnt/data/shekhale/models/nns_graphs)
//   GBDR_RESULTS_ROOT  <root>/<ds>/final_results_<ds>.txt (/home/shekhale/results/nns_graphs)
//   GBDR_LAT_NAME      suffix of the transformed base / graph / net files   (angular_optimal)
//   GBDR_GRAPH_ORIG    original-dimension graph file stem under models/<ds>/ (hnsw_<hnsw_name>); "none" skips the sweep
//   GBDR_GRAPH_LOW     low-dimension graph file stem                         (hnsw_<hnsw_name>_<lat>)
//   GBDR_GRAPH_LOW_NAME  label printed for the low-dimension curve           (hnsw_new_ar); labels that do not
//                      start with "hnsw" get one uniform random entry vertex per query, as in the reference
//   GBDR_SEED          fixes the entry-point generator (default: std::random_device, as the reference)
//   GBDR_NUM_EXPER     repetitions per ef (5)
#include "search_function.h"

static string env_or(const char* name, const string& dflt) {
    const char* s = getenv(name);
    return s && *s ? string(s) : dflt;
}

int main(int argc, char** argv) {
    string datasetName;
    if (argc == 2) {
        datasetName = argv[1];
    } else {
        cout << " Need to specify parameters" << endl;
        return 1;
    }
    cout << datasetName << endl;

    L2Metric l2;
    mt19937 random_gen;
    if (getenv("GBDR_SEED")) {
        random_gen.seed((unsigned)atoll(getenv("GBDR_SEED")));
    } else {
        random_device device;
        random_gen.seed(device());
    }

    const string paramsPath = env_or("GBDR_PARAMS", "/home/shekhale/gbnns_dim_red/search/parameters_of_databases.txt");
    map<string, string> paramsMap = readSearchParams(paramsPath, datasetName);
    const size_t n = atoi(paramsMap["n"].c_str());
    const size_t n_q = atoi(paramsMap["n_q"].c_str());
    const size_t n_tr = atoi(paramsMap["n_tr"].c_str());
    const size_t d = atoi(paramsMap["d"].c_str());
    const size_t d_low = atoi(paramsMap["d_low"].c_str());
    const size_t d_hidden = atoi(paramsMap["d_hidden"].c_str());
    cout << n << " " << n_q << " " << n_tr << " " << d << " " << d_low << endl;
    if (!n || !n_q || !d || !d_low) {
        cout << "dataset " << datasetName << " not described in " << paramsPath << endl;
        return 1;
    }
    vector<int> efs = getVectorFromString(paramsMap["efs"]);
    vector<int> efs_hnsw_origin = getVectorFromString(paramsMap["efs_hnsw"]);
    const string hnsw_name = paramsMap["hnsw_name"];
    const string lat = env_or("GBDR_LAT_NAME", "angular_optimal");

    const string pathData = env_or("GBDR_DATA_ROOT", "/mnt/data/shekhale/data") + "/" + datasetName + "/" + datasetName;
    const string pathModels = env_or("GBDR_MODELS_ROOT", "/mnt/data/shekhale/models/nns_graphs") + "/" + datasetName;

    vector<float> db = loadXvecs<float>(pathData + "_base.fvecs", d, n);
    vector<float> queries = loadXvecs<float>(pathData + "_query.fvecs", d, n_q);
    vector<uint32_t> truth = loadXvecs<uint32_t>(pathData + "_groundtruth.ivecs", n_tr, n_q);
    vector<float> db_ar = loadXvecs<float>(pathData + "_base_" + lat + ".fvecs", d_low, n);

    const string graphOrig = env_or("GBDR_GRAPH_ORIG", "hnsw_" + hnsw_name);
    const string graphLow = env_or("GBDR_GRAPH_LOW", "hnsw_" + hnsw_name + "_" + lat);
    const string lowLabel = env_or("GBDR_GRAPH_LOW_NAME", "hnsw_new_ar");

    const string pathARNets = pathModels + "/" + datasetName + "_net_as_matrix_" + lat;
    const int numberExper = atoi(env_or("GBDR_NUM_EXPER", "5").c_str());
    const int numberThreads = 1;
    Net net;
    net.layerFirst = loadXvecs<float>(pathARNets + "_1.fvecs", d + 1, d_hidden);
    net.layerSecond = loadXvecs<float>(pathARNets + "_2.fvecs", d_hidden + 1, d_hidden);
    net.layerFinal = loadXvecs<float>(pathARNets + "_3.fvecs", d_hidden + 1, d_low);

    const string output_s =
        env_or("GBDR_RESULTS_ROOT", "/home/shekhale/results/nns_graphs") + "/" + datasetName + "/final_results_" + datasetName + ".txt";
    const char* output = output_s.c_str();
    remove(output);

    if (graphOrig != "none") {
        vector<vector<uint32_t>> hnsw = loadEdges(pathModels + "/" + graphOrig + ".ivecs", n, "hnsw");
        performRealTests(n, d, d, n_q, n_tr, efs_hnsw_origin, random_gen, hnsw, hnsw, db, queries, db, queries, truth, output,
                         &l2, "hnsw", false, false, numberExper, numberThreads);
    }
    vector<vector<uint32_t>> hnsw_ar = loadEdges(pathModels + "/" + graphLow + ".ivecs", n, "hnsw_ar");
    performRealNetTests(n, d, d_low, n_q, n_tr, efs, random_gen, hnsw_ar, hnsw_ar, db, queries, db_ar, &net, d_hidden, truth,
                        output, &l2, lowLabel, false, false, numberExper, numberThreads);
    return 0;
}
