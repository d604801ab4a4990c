// gpusimserver without Qt: the reference's command line (main.cpp:12-70) over the C ABI.
//   gpusimserver_b200 [--cpu_only] [--gpu_bitcount N] [--socket PATH] <a.fsim> [b.fsim ...]
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gpusim_b200.h"

static gsb_server* g_server = nullptr;
static void on_signal(int) { if (g_server) gsb_server_stop(g_server); }

int main(int argc, char** argv)
{
    bool cpu_only = false;
    int gpu_bitcount = 0;
    const char* socket_path = "";
    std::vector<const char*> files;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--cpu_only"))
            cpu_only = true;
        else if (!std::strcmp(argv[i], "--gpu_bitcount") && i + 1 < argc)
            gpu_bitcount = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--socket") && i + 1 < argc)
            socket_path = argv[++i];
        else if (!std::strcmp(argv[i], "-h") || !std::strcmp(argv[i], "--help")) {
            std::printf("GPUSim Backend: GPU backend for similarity searching\n"
                        "usage: %s [--cpu_only] [--gpu_bitcount N] [--socket PATH] database.fsim [...]\n", argv[0]);
            return 0;
        } else
            files.push_back(argv[i]);
    }
    if (files.empty()) { // main.cpp:40-43
        std::fprintf(stderr, "Not enough arguments.\n");
        return 1;
    }
    std::fprintf(stderr, "--------------------------\nStarting up GPUSim Server\n--------------------------\n");
    std::fprintf(stderr, "Utilizing %d GPUs for calculation.\n", gsb_device_count());
    if (gsb_server_create(files.data(), static_cast<int>(files.size()), gpu_bitcount, cpu_only ? 0 : 1, &g_server) !=
        GSB_OK) {
        std::fprintf(stderr, "%s\n", gsb_server_last_error());
        return 1;
    }
    if (gsb_server_listen(g_server, socket_path) != GSB_OK) {
        std::fprintf(stderr, "%s\n", gsb_server_last_error());
        return 1;
    }
    std::signal(SIGINT, on_signal);
    std::signal(SIGTERM, on_signal);
    std::fprintf(stderr, "Ready for searches.\n");
    gsb_server_serve(g_server, 0);
    gsb_server_destroy(g_server);
    return 0;
}
