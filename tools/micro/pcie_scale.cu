// pcie_scale.cu -- what caps host<->device bandwidth when several GPUs of one box copy at once?
//
// the engine's end-to-end rate stays flat from 1 to 8 GPUs (per-GPU H2D falls 44 -> 8 GB/s).
// This separates "box ceiling" from "our pinned-memory / process layout" with nothing of the engine in the way:
// for N = 1, 2, 4, 8 devices copying CONCURRENTLY it measures H2D, D2H and both, with
//   process model   one process + one thread per device  |  one process per device (fork before any CUDA call)
//   host memory     cudaHostAlloc | cudaHostAlloc write-combined (H2D source) | mmap 4 KiB pages + cudaHostRegister |
//                   mmap + MADV_HUGEPAGE (THP) + cudaHostRegister | mmap MAP_HUGETLB 2 MiB + cudaHostRegister
//   chunk           48 MB (one 12 MP RGBA image) and 8 MB
// One cudaMemcpyAsync per chunk on a dedicated stream per direction (never the batched-memcpy APIs).
// Output: one JSON document (stdout or --out FILE); profiles/pcie_scaling.json is a committed copy of a run.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pcie_scale pcie_scale.cu -lpthread
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

enum { MAX_DEV = 8, MAX_TESTS = 512 };
enum Alloc { A_HOSTALLOC, A_WC, A_REG4K, A_REGTHP, A_REGHUGETLB, A_COUNT };
static const char *kAllocName[A_COUNT] = {"cudaHostAlloc", "cudaHostAlloc_writeCombined", "mmap4k+cudaHostRegister",
                                          "mmapTHP+cudaHostRegister", "mmapHUGETLB2M+cudaHostRegister"};
enum Dir { D_H2D, D_D2H, D_BOTH, D_COUNT };
static const char *kDirName[D_COUNT] = {"h2d", "d2h", "both"};

struct Test { int alloc, dir, n_dev; size_t chunk; };
struct Result { double t0, t1; double up_bytes, down_bytes; int ok; };

struct Shared {
    pthread_barrier_t bar;
    int n_workers;
    size_t buf_bytes, bytes_per_dir;
    int n_tests;
    Test tests[MAX_TESTS];
    Result res[MAX_TESTS][MAX_DEV];
    int alloc_ok[A_COUNT][MAX_DEV];
    char note[A_COUNT][MAX_DEV][160];
};

static double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct HostBuf { void *p = nullptr; size_t n = 0; int kind = -1; bool registered = false; };

static bool host_alloc(HostBuf &b, int kind, size_t n, char *note)
{
    b.kind = kind;
    b.n = n;
    cudaError_t e = cudaSuccess;
    if (kind == A_HOSTALLOC || kind == A_WC) {
        e = cudaHostAlloc(&b.p, n, cudaHostAllocPortable | (kind == A_WC ? cudaHostAllocWriteCombined : 0));
        if (e != cudaSuccess) { snprintf(note, 160, "cudaHostAlloc: %s", cudaGetErrorString(e)); cudaGetLastError(); b.p = nullptr; return false; }
        return true;
    }
    int flags = MAP_PRIVATE | MAP_ANONYMOUS;
    if (kind == A_REGHUGETLB) flags |= MAP_HUGETLB;
    void *p = mmap(nullptr, n + (2u << 20), PROT_READ | PROT_WRITE, flags, -1, 0);
    if (p == MAP_FAILED) { snprintf(note, 160, "mmap: %s", strerror(errno)); return false; }
    uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1); // 2 MiB aligned start
    if (kind == A_REGHUGETLB) a = (uintptr_t)p;
    if (kind == A_REGTHP && madvise((void *)a, n, MADV_HUGEPAGE) != 0) snprintf(note, 160, "madvise(MADV_HUGEPAGE): %s", strerror(errno));
    if (kind == A_REG4K) madvise((void *)a, n, MADV_NOHUGEPAGE);
    memset((void *)a, 1, n); // fault every page in before pinning
    e = cudaHostRegister((void *)a, n, cudaHostRegisterPortable);
    if (e != cudaSuccess) { snprintf(note, 160, "cudaHostRegister: %s", cudaGetErrorString(e)); cudaGetLastError(); return false; }
    b.p = (void *)a;
    b.registered = true;
    return true;
}

static void host_free(HostBuf &b)
{
    if (!b.p) return;
    if (b.registered) cudaHostUnregister(b.p); // the mapping itself is left to process exit (tests are short-lived)
    else cudaFreeHost(b.p);
    b.p = nullptr;
}

static void worker(int idx, Shared *sh)
{
    cudaSetDevice(idx);
    cudaFree(0);
    void *d_up = nullptr, *d_down = nullptr;
    cudaStream_t s_up, s_down;
    bool dev_ok = cudaMalloc(&d_up, sh->buf_bytes) == cudaSuccess && cudaMalloc(&d_down, sh->buf_bytes) == cudaSuccess;
    cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking);
    int cur_alloc = -1;
    HostBuf h_up, h_down;
    bool host_ok = false;
    for (int t = 0; t < sh->n_tests; t++) {
        const Test T = sh->tests[t];
        if (T.alloc != cur_alloc) {
            host_free(h_up);
            host_free(h_down);
            cur_alloc = T.alloc;
            // write-combined only makes sense as a copy SOURCE: the D2H target of that variant stays plain
            host_ok = dev_ok && host_alloc(h_up, T.alloc, sh->buf_bytes, sh->note[T.alloc][idx]) &&
                      host_alloc(h_down, T.alloc == A_WC ? A_HOSTALLOC : T.alloc, sh->buf_bytes, sh->note[T.alloc][idx]);
            sh->alloc_ok[T.alloc][idx] = host_ok ? 1 : 0;
        }
        Result &R = sh->res[t][idx];
        R = Result{0, 0, 0, 0, 0};
        const bool part = idx < T.n_dev && host_ok;
        // warm-up (untimed): one chunk each way
        if (part) {
            cudaMemcpyAsync(d_up, h_up.p, T.chunk, cudaMemcpyHostToDevice, s_up);
            cudaMemcpyAsync(h_down.p, d_down, T.chunk, cudaMemcpyDeviceToHost, s_down);
            cudaStreamSynchronize(s_up);
            cudaStreamSynchronize(s_down);
        }
        pthread_barrier_wait(&sh->bar);
        if (part) {
            const bool up = T.dir != D_D2H, down = T.dir != D_H2D;
            R.t0 = now_s();
            size_t done = 0;
            while (done < sh->bytes_per_dir) {
                const size_t off = done % (sh->buf_bytes - T.chunk + 1) / 4096 * 4096;
                if (up) cudaMemcpyAsync((char *)d_up + off, (char *)h_up.p + off, T.chunk, cudaMemcpyHostToDevice, s_up);
                if (down) cudaMemcpyAsync((char *)h_down.p + off, (char *)d_down + off, T.chunk, cudaMemcpyDeviceToHost, s_down);
                done += T.chunk;
            }
            cudaError_t e1 = cudaStreamSynchronize(s_up), e2 = cudaStreamSynchronize(s_down);
            R.t1 = now_s();
            R.up_bytes = up ? (double)done : 0;
            R.down_bytes = down ? (double)done : 0;
            R.ok = e1 == cudaSuccess && e2 == cudaSuccess;
        }
        pthread_barrier_wait(&sh->bar);
    }
    host_free(h_up);
    host_free(h_down);
}

static void *thread_main(void *arg)
{
    auto *a = (std::pair<int, Shared *> *)arg;
    worker(a->first, a->second);
    return nullptr;
}

static Shared *make_shared(int n_workers, size_t buf_bytes, size_t bytes_per_dir, const std::vector<Test> &tests)
{
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) { perror("mmap shared"); exit(1); }
    memset(sh, 0, sizeof *sh);
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &at, (unsigned)n_workers);
    sh->n_workers = n_workers;
    sh->buf_bytes = buf_bytes;
    sh->bytes_per_dir = bytes_per_dir;
    sh->n_tests = (int)tests.size();
    for (size_t i = 0; i < tests.size(); i++) sh->tests[i] = tests[i];
    return sh;
}

static std::string read_file(const char *path)
{
    std::string s;
    FILE *f = fopen(path, "r");
    if (!f) return s;
    char buf[256];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    return s;
}

static void emit(FILE *out, const char *model, Shared *sh, bool first_block)
{
    for (int t = 0; t < sh->n_tests; t++) {
        const Test T = sh->tests[t];
        double t0 = 1e300, t1 = 0, up = 0, down = 0, min_dev = 1e300, max_dev = 0;
        int ok = 1, parts = 0;
        for (int i = 0; i < T.n_dev; i++) {
            const Result &R = sh->res[t][i];
            if (!R.ok) { ok = 0; continue; }
            parts++;
            t0 = std::min(t0, R.t0);
            t1 = std::max(t1, R.t1);
            up += R.up_bytes;
            down += R.down_bytes;
            const double per = std::max(R.up_bytes, R.down_bytes) / (R.t1 - R.t0) / 1e9;
            min_dev = std::min(min_dev, per);
            max_dev = std::max(max_dev, per);
        }
        fprintf(out, "%s    {\"model\": \"%s\", \"alloc\": \"%s\", \"dir\": \"%s\", \"n_devices\": %d, \"chunk_mb\": %.0f, \"ok\": %s",
                (first_block && t == 0) ? "" : ",\n", model, kAllocName[T.alloc], kDirName[T.dir], T.n_dev, T.chunk / 1e6,
                (ok && parts == T.n_dev) ? "true" : "false");
        if (ok && parts == T.n_dev && t1 > t0)
            fprintf(out, ", \"h2d_GBps_aggregate\": %.2f, \"d2h_GBps_aggregate\": %.2f, \"per_device_GBps_min\": %.2f, \"per_device_GBps_max\": %.2f",
                    up / (t1 - t0) / 1e9, down / (t1 - t0) / 1e9, min_dev, max_dev);
        else
            fprintf(out, ", \"note\": \"%s\"", sh->note[T.alloc][0]);
        fprintf(out, "}");
    }
}

int main(int argc, char **argv)
{
    int gpus = 0;
    const char *out_path = nullptr;
    size_t buf_mb = 512, total_mb = 3072;
    bool do_threads = true, do_procs = true;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
        else if (!strcmp(argv[i], "--buf-mb") && i + 1 < argc) buf_mb = (size_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--total-mb") && i + 1 < argc) total_mb = (size_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--threads-only")) do_procs = false;
        else if (!strcmp(argv[i], "--procs-only")) do_threads = false;
    }
    if (gpus <= 0) { // count devices without initialising CUDA in this process (children are forked first)
        FILE *p = popen("nvidia-smi -L 2>/dev/null | grep -c '^GPU'", "r");
        if (p) { if (fscanf(p, "%d", &gpus) != 1) gpus = 0; pclose(p); }
    }
    if (gpus <= 0) { fprintf(stderr, "no GPUs visible\n"); return 1; }
    gpus = std::min(gpus, (int)MAX_DEV);

    // MAP_HUGETLB needs a reserved pool: ask for it (root on the box); the variant reports failure otherwise
    const size_t need_pages = (buf_mb * 2 * (size_t)gpus) / 2 + 16;
    const std::string hp_before = read_file("/proc/sys/vm/nr_hugepages");
    {
        FILE *f = fopen("/proc/sys/vm/nr_hugepages", "w");
        if (f) { fprintf(f, "%zu\n", need_pages); fclose(f); }
    }
    const std::string hp_after = read_file("/proc/sys/vm/nr_hugepages");
    const std::string thp = read_file("/sys/kernel/mm/transparent_hugepage/enabled");

    std::vector<int> ns;
    for (int n = 1; n <= gpus; n *= 2) ns.push_back(n);
    if (ns.back() != gpus) ns.push_back(gpus);
    std::vector<Test> tests;
    for (int a = 0; a < A_COUNT; a++)
        for (int n : ns)
            for (int d = 0; d < D_COUNT; d++) {
                if (a == A_WC && d == D_D2H) continue;
                tests.push_back(Test{a, d, n, (size_t)48 << 20});
            }
    for (int n : ns) tests.push_back(Test{A_HOSTALLOC, D_BOTH, n, (size_t)8 << 20}); // smaller DMA chunks
    // group by alloc kind so every worker allocates each kind once
    std::stable_sort(tests.begin(), tests.end(), [](const Test &x, const Test &y) { return x.alloc < y.alloc; });
    if (tests.size() > MAX_TESTS) tests.resize(MAX_TESTS);

    FILE *out = out_path ? fopen(out_path, "w") : stdout;
    if (!out) { perror("open --out"); return 1; }
    char host[128] = "";
    gethostname(host, sizeof host - 1);
    fprintf(out, "{\n  \"tool\": \"tools/micro/pcie_scale.cu\", \"gpus\": %d, \"nproc\": %ld, \"buf_mb_per_direction\": %zu, \"mb_copied_per_direction_per_device\": %zu,\n",
            gpus, sysconf(_SC_NPROCESSORS_ONLN), buf_mb, total_mb);
    fprintf(out, "  \"nr_hugepages_before\": \"%s\", \"nr_hugepages_after_request\": \"%s\", \"thp_enabled\": \"%s\",\n", hp_before.c_str(),
            hp_after.c_str(), thp.c_str());
    fprintf(out, "  \"timing\": \"host CLOCK_MONOTONIC from the first participant's first cudaMemcpyAsync to the last participant's stream sync, all participants released by one barrier\",\n");
    fprintf(out, "  \"results\": [\n");
    bool first = true;

    if (do_procs) { // one process per device; forked BEFORE this process touches CUDA
        Shared *sh = make_shared(gpus, buf_mb << 20, total_mb << 20, tests);
        std::vector<pid_t> kids;
        for (int i = 0; i < gpus; i++) {
            pid_t pid = fork();
            if (pid == 0) { worker(i, sh); _exit(0); }
            kids.push_back(pid);
        }
        for (pid_t k : kids) { int st; waitpid(k, &st, 0); }
        emit(out, "process_per_device", sh, first);
        first = false;
    }
    if (do_threads) { // one process, one thread per device
        Shared *sh = make_shared(gpus, buf_mb << 20, total_mb << 20, tests);
        std::vector<pthread_t> th((size_t)gpus);
        std::vector<std::pair<int, Shared *>> args;
        for (int i = 0; i < gpus; i++) args.push_back({i, sh});
        for (int i = 0; i < gpus; i++) pthread_create(&th[(size_t)i], nullptr, thread_main, &args[(size_t)i]);
        for (int i = 0; i < gpus; i++) pthread_join(th[(size_t)i], nullptr);
        emit(out, "one_process_thread_per_device", sh, first);
        first = false;
    }
    fprintf(out, "\n  ]\n}\n");
    if (out != stdout) fclose(out);
    { // give the hugepage pool back
        FILE *f = fopen("/proc/sys/vm/nr_hugepages", "w");
        if (f) { fprintf(f, "%s\n", hp_before.empty() ? "0" : hp_before.c_str()); fclose(f); }
    }
    return 0;
}
