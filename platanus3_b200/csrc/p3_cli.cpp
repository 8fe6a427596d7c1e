// p3_cli.cpp — `platanus3`-compatible command line (reference main.cpp + src/Options.cpp:23-48 +
// src/ShowInfo.cpp): platanus3_b200 -i {readfile} -k {kmersize} -t {numthread} [-m {filter bits}]
// Writes ./platanus3.log and ./de_bruijn_graph.gfa in the current directory like the reference
// (src/Logging.cpp:11, src/DeBruijnGraph.cpp:454). -d {gpu} selects the device (extension).
#include "../../include/platanus3_b200.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

static void show_usage() { printf("Usage: platanus3 -i {readfile} -k {kmersize} -t {numthread}\n"); }

int main(int argc, char **argv) {
    std::string readfile;
    unsigned long long filter_size = 0;
    int threads = 8, device = 0;          // Options.cpp:12
    unsigned k = 25;                       // Options.cpp:13
    int opt;
    bool parsing = true;
    while (parsing && (opt = getopt(argc, argv, "i:m:k:t:d:")) != -1) {
        switch (opt) {
            case 'i': readfile = optarg; break;
            case 'm': filter_size = strtoull(optarg, nullptr, 10); break;
            case 'k': k = (unsigned)atoi(optarg); break;
            case 't': threads = atoi(optarg); break;
            case 'd': device = atoi(optarg); break;
            default: {
                // Options.cpp:42-45: "Invalid option" goes to the log (appended, Logging.cpp:19), parsing STOPS, and
                // main.cpp:14 carries on with whatever was parsed before it
                FILE *lf = fopen("./platanus3.log", "a");
                if (lf) { fputs("Invalid option\n", lf); fclose(lf); }
                parsing = false;
                break;
            }
        }
    }
    if (readfile.empty()) { show_usage(); return 0; }   // main.cpp:16-19
    unsigned long long st[8];
    int rc = p3_assemble_file(readfile.c_str(), k, filter_size, threads, device, "./de_bruijn_graph.gfa",
                              "./platanus3.log", (uint64_t *)st);
    if (rc != P3_OK) {
        fprintf(stderr, "platanus3_b200: error %d: %s\n", rc, p3_last_error());
        return 1;
    }
    fprintf(stderr, "platanus3_b200: %llu reads, %llu distinct 21-mers, %llu solid %u-mers (+%llu false-positive k-mers), "
                    "%llu junctions, %llu joints, %llu straights\n",
            st[0], st[2], st[3], k, st[4] - st[3], st[5], st[6], st[7]);
    return 0;
}
