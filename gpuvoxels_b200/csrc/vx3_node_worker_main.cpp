// vx3_node_worker — same command line and report as the reference's worker (Voxelyze3/src/Executables/vx3_node_worker.cu:32-142):
//   vx3_node_worker -i <task.vxt> -o <result.vxr> [-f]
// <task.vxt> is the XML the front-end writes ({vxa, input_dir, vxd.f*}, Voxelyze3.cpp:101-111).  The VXD files it lists are
// simulated over its base VXA on all visible GPUs (vx3b_job_run_task) and the report XML is written to <result.vxr>.
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vx3b200.h"

static bool is_file(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
static void usage() {
    printf("This application is called by Voxelyze3, and it can utilize all GPUs on single node to run simulations.\n"
           "Usage: vx3_node_worker -i <task.vxt> -o <result.vxr> [-f] [--devices N] [--fp64]\n"
           "  -h [ --help ]    produce help message\n  -i [ --input ]   Set input file name (.vxt).\n"
           "  -o [ --output ]  Set output file name (.vxr).\n  -f [ --force ]   Overwrite output file if exists.\n");
}

int main(int argc, char **argv) {
    std::string input, output;
    bool force = false;
    int devices = 0, precision = VX3B_FP32;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-i" || a == "--input") && i + 1 < argc)
            input = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc)
            output = argv[++i];
        else if (a == "-f" || a == "--force")
            force = true;
        else if (a == "--devices" && i + 1 < argc)
            devices = atoi(argv[++i]);
        else if (a == "--fp64")
            precision = VX3B_FP64;
        else {
            usage();
            return 1;
        }
    }
    if (input.empty() || output.empty()) {
        usage();
        return 1;
    }
    if (!is_file(input)) {
        printf("Input file not found.\n\n");
        usage();
        return 1;
    }
    if (is_file(output) && !force) {
        printf("Output file already exist.\n\n");
        usage();
        return 1;
    }
    if (devices <= 0) {
        const char *e = getenv("VX3B_DEVICES");
        devices = e ? atoi(e) : 0;
    }
    if (devices <= 0) { // all visible GPUs: probe by creating contexts until one fails
        devices = 0;
        for (int d = 0; d < 64; d++) {
            vx3b_ctx *c = nullptr;
            if (vx3b_ctx_create(d, precision, &c) != VX3B_OK)
                break;
            vx3b_ctx_destroy(c);
            devices++;
        }
        if (devices == 0) {
            printf("ERROR: %s\n", vx3b_last_error());
            return 1;
        }
    }
    // worlds that ask for it (VXA.Simulator.RecordHistory.RecordStepSize) stream their .history to stdout first, as the reference's
    // device-side printf does (VX3_SimulationManager.cu:40-114); `Voxelyze3 -i <dir> > a.history` captures it
    int rc = vx3b_job_run_task(input.c_str(), output.c_str(), devices, precision, 1);
    if (rc != VX3B_OK) {
        printf("ERROR: %s\n", vx3b_last_error());
        return 1;
    }
    return 0;
}
