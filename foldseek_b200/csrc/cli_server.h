// cli_server.h - gpuserver: module calls over a Unix-domain socket, client side of --gpu-server 1
// Part of the host module (foldseek-b200), included by cli_main.cpp inside its anonymous namespace.
#ifndef CLI_SERVER_H
#define CLI_SERVER_H

// ---- gpuserver: resident target DB, module calls over a Unix-domain socket ----
// GPUSharedMemory::getShmHash (lib/mmseqs/src/commons/GpuUtil.cpp:18-34): real path of the DB + CUDA_VISIBLE_DEVICES + version
std::string server_socket_path(const std::string &targetDb) {
    std::string id = real_path(targetDb);
    if (const char *v = getenv("CUDA_VISIBLE_DEVICES")) id += v;
    id += fsb200_build_id();
    uint64_t h = 1469598103934665603ull;   // FNV-1a
    for (unsigned char c : id) { h ^= c; h *= 1099511628211ull; }
    const char *dir = getenv("FSB200_SOCKET_DIR");
    char name[64]; snprintf(name, sizeof(name), "/fsb200-gpuserver-%016llx.sock", (unsigned long long) h);
    return std::string(dir ? dir : "/tmp") + name;
}

bool send_all(int fd, const void *buf, size_t n) {
    const char *p = static_cast<const char *>(buf);
    while (n > 0) { const ssize_t w = send(fd, p, n, MSG_NOSIGNAL); if (w <= 0) return false; p += w; n -= (size_t) w; }
    return true;
}
bool recv_all(int fd, void *buf, size_t n) {
    char *p = static_cast<char *>(buf);
    while (n > 0) { const ssize_t r = recv(fd, p, n, 0); if (r <= 0) return false; p += r; n -= (size_t) r; }
    return true;
}
bool send_str(int fd, const std::string &t) { const uint32_t n = (uint32_t) t.size(); return send_all(fd, &n, 4) && send_all(fd, t.data(), n); }
bool recv_str(int fd, std::string &t) {
    uint32_t n = 0;
    if (!recv_all(fd, &n, 4) || n > (1u << 28)) return false;
    t.resize(n);
    return n == 0 || recv_all(fd, &t[0], n);
}

int connect_to(const std::string &path) {
    const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    sockaddr_un addr; memset(&addr, 0, sizeof(addr)); addr.sun_family = AF_UNIX;
    if (path.size() >= sizeof(addr.sun_path)) { close(fd); die("socket path too long: " + path); }
    strcpy(addr.sun_path, path.c_str());
    if (connect(fd, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0) { close(fd); return -1; }
    return fd;
}

// client side of --gpu-server 1 (ungappedprefilter.cpp:72-124: wait for the server of the target DB, then use it).  The prefilter's
// target is <targetDB>_ss while the server is started on <targetDB>: both names lead to the same server.
int forward_to_server(int argc, char **argv, const std::string &targetPath, const Args &a) {
    std::vector<std::string> names(1, targetPath);
    if (targetPath.size() > 3 && targetPath.compare(targetPath.size() - 3, 3, "_ss") == 0) names.push_back(targetPath.substr(0, targetPath.size() - 3));
    int fd = -1;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool told = false;
    for (;;) {
        for (const std::string &n : names) { fd = connect_to(server_socket_path(n)); if (fd >= 0) break; }
        if (fd >= 0) break;
        const long elapsed = (long) std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
        if (a.gpuServerWait == 0 || (a.gpuServerWait > 0 && elapsed >= a.gpuServerWait))
            die("gpuserver for database " + targetPath + " not found" + (a.gpuServerWait > 0 ? " after " + std::to_string(elapsed) + " seconds" : std::string()) +
                ".\nPlease start `foldseek-b200 gpuserver` on it with the same CUDA_VISIBLE_DEVICES");
        if (!told) { say_err("Waiting for `gpuserver`\n"); told = true; }
        std::this_thread::sleep_for(std::chrono::milliseconds(200));
    }
    char cwd[PATH_MAX];
    if (getcwd(cwd, sizeof(cwd)) == NULL) die("getcwd failed");
    bool ok = send_str(fd, cwd);
    const uint32_t n = (uint32_t) argc;
    ok = ok && send_all(fd, &n, 4);
    for (int i = 0; ok && i < argc; i++) ok = send_str(fd, argv[i]);
    int32_t code = EXIT_FAILURE; std::string err, out;
    ok = ok && recv_all(fd, &code, 4) && recv_str(fd, err) && recv_str(fd, out);
    close(fd);
    if (!ok) die("connection to the gpuserver for database " + targetPath + " broke");
    fputs(out.c_str(), stdout); fputs(err.c_str(), stderr);
    return code;
}

volatile sig_atomic_t g_keepRunning = 1;
void stop_handler(int) { g_keepRunning = 0; }

int run_module(int argc, char **argv);

int gpuserver_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 1) die("usage: foldseek-b200 gpuserver <i:targetDB> [options]   (serves prefilter / structurealign calls that carry --gpu-server 1)");
    const std::string sock = server_socket_path(a.pos[0]);
    { const int probe = connect_to(sock); if (probe >= 0) { close(probe); die("a gpuserver for database " + a.pos[0] + " is already running (" + sock + ")"); } }
    unlink(sock.c_str());   // left behind by a server that was killed
    Resident resident;
    resident.gpus = a.gpus; resident.lanes = a.lanes; resident.threads = a.threads;
    g_resident = &resident;
    // Preload with the parameters of this command line (defaults where none are given): a later call whose parameters differ in a way
    // that matters for the staged state replaces the state (the staging is then paid once more).  --comp-bias-corr-scale is per module
    // in the workflow (0.15 prefilter, 0.5 structurealign, StructureSearch.cpp:101-107), the defaults of fsb200_config.
    const bool haveSs = !slurp(a.pos[0] + "_ss.dbtype", false).empty();
    const bool haveCa = !slurp(a.pos[0] + "_ca.dbtype", false).empty();
    if (haveSs) {
        pref_target(a, argv[0], a.pos[0] + "_ss", (size_t) -1);
        aln_target(a, argv[0], a.pos[0], (size_t) -1, haveCa, true);   // with the index: serves structurealign and search calls
    } else say_err("gpuserver: " + a.pos[0] + "_ss not found, nothing preloaded\n");
    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) die("socket() failed");
    sockaddr_un addr; memset(&addr, 0, sizeof(addr)); addr.sun_family = AF_UNIX;
    if (sock.size() >= sizeof(addr.sun_path)) die("socket path too long: " + sock);
    strcpy(addr.sun_path, sock.c_str());
    if (bind(ls, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0 || listen(ls, 16) != 0) die("cannot listen on " + sock);
    struct sigaction act; memset(&act, 0, sizeof(act)); act.sa_handler = stop_handler;
    sigaction(SIGINT, &act, NULL); sigaction(SIGTERM, &act, NULL);
    say_out(sock + "\n");   // the reference prints the name of its shared-memory block once the DB is loaded (gpuserver.cpp:81)
    while (g_keepRunning) {
        pollfd pf{ls, POLLIN, 0};
        if (poll(&pf, 1, 200) <= 0) continue;
        const int fd = accept(ls, NULL, NULL);
        if (fd < 0) continue;
        { timeval tv{30, 0}; setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv)); }   // a caller that connects and stays silent must not block the others
        std::string cwd; uint32_t n = 0; std::vector<std::string> args;
        bool ok = recv_str(fd, cwd) && recv_all(fd, &n, 4) && n >= 2 && n < 4096;
        for (uint32_t i = 0; ok && i < n; i++) { std::string t; ok = recv_str(fd, t); args.push_back(t); }
        if (!ok) { close(fd); continue; }
        int32_t code = EXIT_FAILURE;
        { std::lock_guard<std::mutex> g(g_console.m); g_console.capture = true; g_console.out.clear(); g_console.err.clear(); }
        resident.touched = false;
        try {
            if (chdir(cwd.c_str()) != 0) die("gpuserver: cannot change to " + cwd);
            if (args[1] != "prefilter" && args[1] != "structurealign" && args[1] != "search") die("gpuserver: module " + args[1] + " is not served");
            std::vector<char *> av; for (std::string &t : args) av.push_back(&t[0]);
            av[0] = argv[0];
            code = run_module((int) av.size(), av.data());
        } catch (const std::exception &e) {
            say_err(std::string("Error: ") + e.what() + "\n");
            code = EXIT_FAILURE;
            if (resident.touched) { resident.pref.reset(); resident.aln.reset(); }   // a context may be mid-batch: restage at the next call
        }
        std::string err, out;
        { std::lock_guard<std::mutex> g(g_console.m); g_console.capture = false; err.swap(g_console.err); out.swap(g_console.out); }
        if (!(send_all(fd, &code, 4) && send_str(fd, err) && send_str(fd, out))) say_err("gpuserver: client went away\n");
        close(fd);
    }
    close(ls);
    unlink(sock.c_str());
    g_resident = NULL;
    return EXIT_SUCCESS;
}

#endif
