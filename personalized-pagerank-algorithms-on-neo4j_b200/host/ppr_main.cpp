// ppr_main.cpp -- `PPR`, the reference's command-line driver (J/PPR.java:154-200), over libpprb200.so.
// Same options and defaults: -alpha 0.15 -eps 0.5 -query 50 -k 10 -node name -label Person -rel Relation -db <dir> -help.
// -db names a Neo4j 3.x store directory (target/got.db: nodestore / relationshipstore records are read natively, no JVM) or a
// directory with the Neo4j import CSVs (*_Nodes.csv, *_Rels.csv); -out is where
// <db>_AlgoPerfResults.txt is appended (default: current directory).
#include <dirent.h>

#include <cstring>
#include <iostream>

#include "ppr_host.hpp"

static std::string find_csv(const std::string& dir, const char* suffix) {
    DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("cannot open directory " + dir);
    std::string best;
    while (dirent* e = readdir(d)) {
        std::string f = e->d_name;
        if (f.size() >= strlen(suffix) && f.compare(f.size() - strlen(suffix), strlen(suffix), suffix) == 0 && (best.empty() || f < best)) best = f;
    }
    closedir(d);
    if (best.empty()) throw std::runtime_error(std::string("no *") + suffix + " and no neostore.nodestore.db under " + dir);
    return dir + "/" + best;
}

int main(int argc, char** argv) {
    double alpha = 0.15, eps = 0.5;
    int query = 50, k = 10;
    std::string node = "name", label = "Person", rel = "Relation", db = "tests/golden/got", out = ".";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value for " + a); return argv[++i]; };
        if (a == "-help" || a == "--help") {
            printf("usage: PPR\n -alpha <arg>  The possibility that a random walk stops at current node (Default: 0.15)\n -eps <arg>    The relative error bound (Default: 0.5)\n"
                   " -query <arg>  The number of queries for the test (Default: 50)\n -k <arg>      For Top-k Algorithm tests (Default: 10)\n -node <arg>   The node property (Default: \"name\")\n"
                   " -label <arg>  The nodes' label type (Default: \"Person\")\n -rel <arg>    The relationships' type (Default: \"Relation\")\n -db <arg>     The directory of the input database (Neo4j store or import CSVs)\n -help\n");
            return 0;
        } else if (a == "-alpha" || a == "--alpha") alpha = std::stod(val());
        else if (a == "-eps" || a == "--epsilon") eps = std::stod(val());
        else if (a == "-query" || a == "--query") query = std::stoi(val());
        else if (a == "-k" || a == "--k") k = std::stoi(val());
        else if (a == "-node" || a == "--node_property") node = val();
        else if (a == "-label" || a == "--label_type") label = val();
        else if (a == "-rel" || a == "--rel_type") rel = val();
        else if (a == "-db" || a == "--db_dir") db = val();
        else if (a == "-out") out = val();
        else { fprintf(stderr, "Unrecognized option: %s\n", a.c_str()); return 2; }
    }
    fora_neo4j::PPR t;
    try {
        t.createDb(db);
        if (std::ifstream(db + "/neostore.nodestore.db").good()) t.setupAdjMatrix(db);   // a Neo4j 3.x store directory (the reference's -db), read natively
        else t.setupAdjMatrix(find_csv(db, "Nodes.csv"), find_csv(db, "Rels.csv"));
        t.set_configuration(alpha, eps, node, label, rel);
        printf("\nCreating index ...\nIndex created\n");
        t.algo_perf_batch_test(query, k, out);
        for (auto& r : t.rows)
            printf("%-18s param %-10s avg %.3f ms  max_err %.3e  precision %.3f  NDCG %.4f\n", r.algo.c_str(), fora_neo4j::java_double(r.param).c_str(), r.avg_ms, r.max_err, r.precision, r.ndcg);
        t.shutDown();
    } catch (const std::exception& e) {   // J/PPR.java:196-199
        printf("Algo performance batch test failed!\n%s\n", e.what());
        return 1;
    }
    return 0;
}
