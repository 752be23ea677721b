// Voxelyze3 — same command line as the reference's front-end (Voxelyze3/src/Executables/Voxelyze3.cpp:25-131):
//   Voxelyze3 -i <dir> [-o <report.xml>] [-l] [-f] [-w <worker>] [-n N]
// It validates the arguments exactly as the reference does (return code 1 on usage errors), writes the task file
// workspace/locally/<timestamp>.vxt, runs the worker on it and copies the worker's .vxr to -o.  Return code 0 even if
// the worker failed (the reference only prints "Worker failed to finish the job").
#include <dirent.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <string>
#include <vector>

static bool is_file(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
static bool is_dir(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
static std::string xml_escape(const std::string &s) {
    std::string o;
    for (char c : s)
        o += c == '&' ? "&amp;" : c == '<' ? "&lt;" : c == '>' ? "&gt;" : std::string(1, c);
    return o;
}
static void usage() {
    printf("Thank you for using Voxelyze3. This program should be run on a computer that has GPUs.\n"
           "Usage:\n\nVoxelyze3 -i <data_path> -o <report_path> -l -f\n\nAllowed options:\n"
           "  -h [ --help ]         produce help message\n  -l [ --locally ]      If this machine already has GPUs, locally run tasks on this machine.\n"
           "  -n [ --nodes ] arg    The number of nodes you want to start simultaneously. (Not implemented.)\n"
           "  -i [ --input ] arg    Set input directory path which contains a generation of VXA files.\n"
           "  -o [ --output ] arg   Set output file path for report. (e.g. report_1.xml)\n"
           "  -w [ --worker ] arg   Specify which worker you want to use. vx3_node_worker by default.\n"
           "  -f [ --force ]        Overwrite output file if exists.\n\n");
}

int main(int argc, char **argv) {
    std::string input, output, worker = "./vx3_node_worker";
    bool force = false, locally = false, nodes = false, help = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-i" || a == "--input") && i + 1 < argc)
            input = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc)
            output = argv[++i];
        else if ((a == "-w" || a == "--worker") && i + 1 < argc)
            worker = argv[++i];
        else if ((a == "-n" || a == "--nodes") && i + 1 < argc) {
            nodes = true;
            ++i;
        } else if (a == "-l" || a == "--locally")
            locally = true;
        else if (a == "-f" || a == "--force")
            force = true;
        else if (a == "-lf" || a == "-fl")
            locally = force = true;
        else
            help = true;
    }
    if (help || input.empty()) {
        usage();
        return 1;
    }
    if (nodes && locally) {
        printf("ERROR: -n and -s cannot use together.\n");
        usage();
        return 1;
    }
    if (!nodes)
        printf("Running simulation locally by default.\n");
    if (!output.empty() && is_file(output) && !force) {
        printf("Error: output file exists.\n\n");
        usage();
        return 1;
    }
    if (!is_dir(input)) {
        printf("Error: input directory not found.\n\n");
        usage();
        return 1;
    }
    if (!is_file(input + "/base.vxa")) {
        printf("No base.vxa found in input directory.\n\n");
        usage();
        return 1;
    }
    if (!is_file(worker)) { // also accept the worker that sits next to this binary
        std::string self = argv[0];
        size_t slash = self.find_last_of('/');
        std::string beside = (slash == std::string::npos ? std::string(".") : self.substr(0, slash)) + "/vx3_node_worker";
        if (worker == "./vx3_node_worker" && is_file(beside))
            worker = beside;
        else {
            printf("Need an executable worker but nothing found.\n\n");
            usage();
            return 1;
        }
    }
    if (nodes)
        return 0; // "Not implemented." in the reference as well (empty branch, Voxelyze3.cpp:128-130)
    mkdir("workspace", 0777);
    mkdir("workspace/locally", 0777);
    char stamp[32];
    time_t now = time(nullptr);
    strftime(stamp, sizeof(stamp), "%Y%m%d%H%M%S", localtime(&now));
    std::string vxt = std::string("workspace/locally/") + stamp + "_" + std::to_string(getpid()) + ".vxt";
    std::string vxr = vxt.substr(0, vxt.size() - 4) + ".vxr";
    std::vector<std::string> files;
    if (DIR *d = opendir(input.c_str())) {
        while (dirent *e = readdir(d)) {
            std::string n = e->d_name;
            if (n.size() > 4) {
                std::string ext = n.substr(n.size() - 4);
                std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char)tolower(c); });
                if (ext == ".vxd")
                    files.push_back(n);
            }
        }
        closedir(d);
    }
    std::sort(files.begin(), files.end());
    {
        std::ofstream t(vxt);
        t << "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<vxa>" << xml_escape(input + "/base.vxa") << "</vxa><input_dir>" << xml_escape(input) << "</input_dir><vxd>";
        for (auto &f : files)
            t << "<f>" << xml_escape(f) << "</f>";
        t << "</vxd>\n";
    }
    std::string command = worker + " -i " + vxt + " -o " + vxr;
    printf("%s\n", command.c_str());
    fflush(stdout);
    pid_t pid = fork();
    if (pid == 0) {
        execl(worker.c_str(), worker.c_str(), "-i", vxt.c_str(), "-o", vxr.c_str(), (char *)nullptr);
        _exit(127);
    }
    int st = 0;
    if (pid > 0)
        waitpid(pid, &st, 0);
    if (is_file(vxr)) {
        if (!output.empty()) {
            std::ifstream src(vxr, std::ios::binary);
            std::ofstream dst(output, std::ios::binary | std::ios::trunc);
            dst << src.rdbuf();
            if (!dst)
                printf("ERROR: Failed to copy result file: %s.\n", vxr.c_str());
        }
    } else
        printf("File not exist: %s. Worker failed to finish the job.\n", vxr.c_str());
    return 0;
}
