// global_faldoi -- drop-in for the reference executable (src/global_faldoi.cpp:1846-2213):
//
//   global_faldoi ims.txt in_flow.flo out.flo [occ_in.png occ_out.png]
//                 [-m method] [-w warps] [-p params.txt] [-glb_iters n] [-verbose 0|1]
//
// Same argv contract, file formats, method ids, parameter defaults and messages; the
// minimisation itself (tvl2OF / nltvl1_PD / tvcsad_PD / nltvcsad_PD /
// guided_tvl2coupled_occ) runs on B200s through the C ABI in include/faldoi_gpu.h.
// Extra options (unknown to the reference's scripts):
//   -device d        CUDA device of a single call (default 0)
//   -devices a,b,..  the GPUs to use ("all" = every visible one).  A single call on a frame of 4K or more
//                    (w*h >= 3840*2160, TVL2) is cut into row stripes over them (halo rows exchanged over NVLink);
//                    a sequence is sharded by pair, one host thread + batched handle per GPU
//   -stripes 0       never cut a frame into stripes
//   -seq_stats 1     per GPU, where the host thread of a sequence spent its time (stderr)
//   -pinned 0|1      sequence staging buffers in ordinary / page-locked memory (default: page-locked for long sequences)
//   -host_preproc 1  run main()'s preprocessing on the host instead of the GPU
//   -seq jobs.txt    many pairs in one process (one job per line: the positional arguments of a normal call).
//                    Files are read and decoded by a pool of host threads, jobs of equal size fill the slots of a
//                    batched solver handle (-batch n, default 16) through pinned staging buffers, results are
//                    written by background threads -- SURVEY 8f rows 3 and 4; what the reference's scripts do with
//                    one process per pair (scripts_python/faldoi_sift.py:314-318)
// There is no CPU fallback: without a usable GPU the program reports the error and fails.
#include <malloc.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/faldoi_gpu.h"
#include "image_io.h"
#include "preprocess.h"

using faldoi_host::Image;

namespace {

// pick_option (src/utils_preprocess.cpp:21-35): "-name value" anywhere on the line,
// removed from the vector; a trailing "-name" without a value is left in place.
std::string pick_option(std::vector<std::string> &args, const std::string &option, const std::string &def) {
    const std::string flag = "-" + option;
    for (auto it = args.begin(); it != args.end(); ++it) {
        if (*it == flag) {
            if (it + 1 == args.end()) continue;
            const std::string value = *(it + 1);
            args.erase(it, it + 2);
            return value;
        }
    }
    return def;
}

void print_today() {
    const std::time_t tt = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
    std::cerr << "today is: " << std::ctime(&tt);
}

void usage(size_t n) {
    fprintf(stderr, "Without occlusions:\n");
    fprintf(stderr, "Usage: %lu  ims.txt in_flow.flo  out.flo "
                    "[-m method_val] [-w num_warps] [-p file of parameters] [-glb_iters global_iters] [-verbose verbose]"
                    " \n", n);
    fprintf(stderr, "With occlusions:\n");
    fprintf(stderr, "Usage: %lu  ims.txt in_flow.flo  out.flo occl_input.png occl_out.png"
                    " [-m method_val] [-w num_warps] [-p file of parameters] [-glb_iters global_iters] [-verbose verbose]"
                    "\n", n);
}

const char *method_name(int m) {
    switch (m) {
        case FALDOI_M_NLTVL1: return "NLTV-L1";
        case FALDOI_M_TVCSAD: return "TV-CSAD";
        case FALDOI_M_NLTVCSAD: return "NLTV-CSAD";
        case FALDOI_M_TVL1_W: return "TV-l2 coupled Weights";
        case FALDOI_M_NLTVCSAD_W: return "NLTV-CSAD Weights";
        case FALDOI_M_NLTVL1_W: return " NLTV-L1 Weights";
        case FALDOI_M_TVCSAD_W: return "TV-CSAD Weights";
        default: return "TV-l2 coupled";
    }
}

bool is_nltv(int m) {
    return m == FALDOI_M_NLTVL1 || m == FALDOI_M_NLTVL1_W || m == FALDOI_M_NLTVCSAD || m == FALDOI_M_NLTVCSAD_W;
}

struct Options {
    int val_method = 0, nwarps = 5, glb_it = 400, device = 0, batch = 16, pinned = -1;
    bool verbose = false, host_preproc = false, stripes = true, seq_stats = false;
    std::vector<int> devices;
    std::string file_params;
};

// Page-locked staging memory of one job in flight (sequence mode): the loaders decode the frames, the flow and the
// occlusion mask straight into it, the GPU copies from and back into it, the writers save from it.
struct Slot {
    float *pin = nullptr;
    size_t cap = 0;  // floats
};

// Staging memory for the slots, carved out of a few large arenas that live as long as the sequence (slots are
// reused from job to job, so no job pays for fresh pages).  Page-locked arenas make the copies to and from the GPU
// asynchronous and faster, but page-locking is slow -- measured on the B200 hosts: 40 ms per 22 MB slot, and
// concurrent cudaHostAlloc calls serialise inside the driver -- against ~2 ms per pair for staged copies from
// ordinary memory; it pays off once every slot has been reused some twenty times.  Hence: page-locked for long
// sequences (or -pinned 1), ordinary memory otherwise (or -pinned 0); one thread allocates at a time, a few slots
// per call, and only as the pipeline actually fills.
class PinnedArenas {
  public:
    PinnedArenas(size_t slots_per_arena, bool pinned) : per_arena_(slots_per_arena ? slots_per_arena : 1), pinned_(pinned) {}
    ~PinnedArenas() {
        for (float *a : arenas_) pinned_ ? faldoi_pinned_free(a) : free(a);
    }
    // memory for one slot of `need` floats (nullptr on failure: the job then decodes into its own vectors)
    float *carve(size_t need) {
        std::lock_guard<std::mutex> l(mu_);
        if (need != slot_floats_ || left_ == 0) {
            float *a = nullptr;
            if (pinned_) {
                a = (float *)faldoi_pinned_alloc(need * per_arena_ * sizeof(float));
            } else if (posix_memalign((void **)&a, 4096, need * per_arena_ * sizeof(float)) != 0) {
                a = nullptr;
            }
            if (!a) return nullptr;
            arenas_.push_back(a);
            cur_ = a, slot_floats_ = need, left_ = per_arena_;
        }
        float *p = cur_;
        cur_ += need;
        left_--;
        return p;
    }

  private:
    std::mutex mu_;
    std::vector<float *> arenas_;
    float *cur_ = nullptr;
    size_t per_arena_, slot_floats_ = 0, left_ = 0;
    bool pinned_;
};
PinnedArenas *g_arenas = nullptr;
bool g_have_gpu = true;  // false when no CUDA device is visible (page-locking needs the driver)
// -seq_stats 1: where the host side of a sequence spends its time (microseconds, summed over all threads)
std::atomic<long long> g_us_pin{0}, g_us_decode{0}, g_us_save{0}, g_us_prepare{0}, g_us_wait_load{0};
struct UsTimer {
    std::atomic<long long> &acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit UsTimer(std::atomic<long long> &a) : acc(a) {}
    ~UsTimer() { acc += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

// Stage 1 of a job (host I/O, no messages): ims.txt + the frames, the flow and the occlusion
// mask.  A single call decodes its files concurrently; a sequence decodes many jobs at once
// instead.  Errors are kept for the solve stage to report in job order.
struct Loaded {
    Image i_1, i0, i1, flow, occ;
    int num_files = 0;
    std::exception_ptr err;
};

Loaded load_inputs(const std::vector<std::string> &args, int val_method, bool parallel, Slot *slot = nullptr) {
    Loaded L;
    try {
        // ims.txt: line 1 = I0, line 2 = I1, line 3 = I-1, line 4 = I2 (unused)
        std::string filename_i_1, filename_i0, filename_i1, line;
        {
            std::ifstream infile(args[1]);
            while (std::getline(infile, line)) {
                ++L.num_files;
                if (L.num_files == 1) filename_i0 = line;
                if (L.num_files == 2) filename_i1 = line;
                if (L.num_files == 3) filename_i_1 = line;
            }
        }
        // with fewer than 4 lines the reference reads I1 in place of I-1 (:1933-1937)
        const std::string third = (L.num_files == 4) ? filename_i_1 : filename_i1;
        const bool third_is_i1 = (third == filename_i1);  // the same file: decode it once
        // staging layout from the first frame's header: three frames of pd planes, the flow (2), the mask (1)
        faldoi_host::Sink s0, s1, s2, sf, so;
        if (slot) {
            try {
                int w = 0, h = 0, pd = 0;
                faldoi_host::probe_image(filename_i0, &w, &h, &pd);
                const size_t n = (size_t)w * h, need = (3 * (size_t)pd + 3) * n;
                if (need > slot->cap && g_arenas) {
                    UsTimer tp(g_us_pin);
                    slot->pin = g_arenas->carve(need);
                    slot->cap = slot->pin ? need : 0;
                }
                if (slot->pin) {
                    float *p = slot->pin;
                    s0 = {p, pd * n}, s1 = {p + pd * n, pd * n}, s2 = {p + 2 * pd * n, pd * n};
                    sf = {p + 3 * pd * n, 2 * n}, so = {p + 3 * pd * n + 2 * n, n};
                }
            } catch (...) {  // the real read below reports what is wrong with the file
            }
        }
        UsTimer td(g_us_decode);
        auto rd = [](std::string f, faldoi_host::Sink s) { return faldoi_host::read_image_split(f, s); };
        const auto policy = parallel ? std::launch::async : std::launch::deferred;
        std::future<Image> f0 = std::async(policy, rd, filename_i0, s0);
        std::future<Image> f1 = std::async(policy, rd, filename_i1, s1);
        std::future<Image> ff = std::async(policy, rd, args[2], sf);
        std::future<Image> fo;
        if (val_method >= 8) fo = std::async(policy, rd, args.size() == 6 ? args[4] : std::string(), so);
        std::exception_ptr first;
        auto take = [&](std::future<Image> &f, Image &dst) {
            try {
                dst = f.get();
            } catch (...) {
                if (!first) first = std::current_exception();
            }
        };
        try {  // same order as the reference reads them, so the same file is blamed first
            if (!third_is_i1) L.i_1 = rd(third, s2);
        } catch (...) {
            first = std::current_exception();
        }
        take(f0, L.i0);
        take(f1, L.i1);
        if (third_is_i1) L.i_1 = L.i1;
        take(ff, L.flow);
        if (val_method >= 8) take(fo, L.occ);
        L.err = first;
    } catch (...) {
        L.err = std::current_exception();
    }
    return L;
}

// A job after main()'s checks (:1949-2022): what the solver needs and where the results go.
struct Ready {
    Loaded in;
    int method = 0, w = 0, h = 0, pd = 0;
    bool passthrough = false;  // method id outside 0..8: the reference writes the input flow back unchanged
    faldoi_params params{};
    float *u = nullptr, *chi = nullptr;  // flow (u1 | u2) and occlusion mask, in and out: the decoded images' storage
    int slot = -1;                       // staging slot to give back once the results are on disk (sequence mode)
    std::string flow_file, occ_file;
    faldoi_log log{};
    double secs = 0;
};

std::mutex g_io_mu;  // whole messages of concurrent jobs do not interleave

// The checks and messages of main() between reading the files and calling the solver.  Returns EXIT_SUCCESS or
// the reference's failure code; throws what the loaders threw.
int prepare_job(const std::vector<std::string> &args, const Options &opt, Loaded &&loaded, Ready &R) {
    R.in = std::move(loaded);
    if (R.in.err) std::rethrow_exception(R.in.err);
    const Image &i_1 = R.in.i_1, &i0 = R.in.i0, &i1 = R.in.i1, &flow = R.in.flow, &occ = R.in.occ;
    int val_method = opt.val_method;
    const int num_files = R.in.num_files;
    R.flow_file = args[3];
    if (args.size() == 6) R.occ_file = args[5];

    auto same = [](const Image &a, const Image &b) { return a.w == b.w && a.h == b.h && a.pd == b.pd; };
    if (num_files == 3) {
        if (!same(i0, i1) || !same(i0, i_1) || !same(i1, i_1)) return fprintf(stderr, "ERROR: input images and flow size mismatch\n");
    } else if (!same(i0, i1)) {
        return fprintf(stderr, "ERROR: input images and flow size mismatch\n");
    }
    if (i0.w != flow.w || i0.h != flow.h || flow.pd != 2) return fprintf(stderr, "ERROR: input flow field size mismatch\n");

    if (num_files == 2 && val_method == FALDOI_M_TVL1_OCC) {
        fprintf(stderr, "Since only two images given, method is changed to TV-l2 coupled\n");
        fprintf(stderr, "Occlusion estimation requires 4 frames: i_1 ==> i0 ==> i1 ==> i2\n");
        val_method = FALDOI_M_TVL1;
    } else if (num_files == 4 && val_method >= 0 && val_method <= 7) {
        fprintf(stderr, "Only two of the four images given will be used, according to the method selected\n");
        fprintf(stderr, "Method: %s\n", method_name(val_method));
    } else {
        fprintf(stderr, "Method: ");
        if (val_method == FALDOI_M_TVL1_OCC) fprintf(stderr, "TV-l2 occlusions\n");
    }
    // Method ids outside 0..8: the reference's dispatch (:2132-2167) matches nothing and main() writes the
    // input flow back unchanged -- so do we (after the same file checks), without touching the GPU.
    R.passthrough = (val_method < 0 || val_method > 8);
    R.method = val_method;
    R.w = i0.w, R.h = i0.h, R.pd = i0.pd;
    const size_t size = (size_t)R.w * R.h;
    if (val_method == FALDOI_M_TVL1_OCC && (occ.w != R.w || occ.h != R.h)) return fprintf(stderr, "ERROR: input images and flow size mismatch\n");

    if (faldoi_params_from_file(opt.file_params.c_str(), R.passthrough ? 0 : val_method, opt.glb_it, &R.params) != FALDOI_OK) {
        fprintf(stderr, "ERROR: %s\n", faldoi_last_error());
        return EXIT_FAILURE;
    }
    R.params.warps = opt.nwarps;
    if (opt.verbose)
        std::cerr << "Parameters: \n lambda: " << R.params.lambda << ", theta: " << R.params.theta << ", beta: " << R.params.beta
                  << ", alpha: " << R.params.alpha << ", \n tau_u: " << R.params.tau_u << ", tau_eta: " << R.params.tau_eta
                  << ", tau_chi: " << R.params.tau_chi << ", mu: " << R.params.mu << "\n";
    if (is_nltv(val_method)) {
        std::printf("W:%d H:%d Pd:%d\n", R.w, R.h, R.pd);
        if (R.pd < 3) {
            fprintf(stderr, "ERROR: the NLTV models need a colour (3-channel) first frame\n");
            return EXIT_FAILURE;
        }
    }
    if (R.w < 5 || R.h < 5) {  // gaussian() aborts with "sigma too large" on such frames
        fprintf(stderr, "GaussianSmooth: sigma too large\n");
        return EXIT_FAILURE;
    }
    (void)size;
    R.u = R.in.flow.px();  // u1 | u2
    if (val_method >= 8 && !R.passthrough) R.chi = R.in.occ.px();
    if (val_method == FALDOI_M_NLTVL1 || val_method == FALDOI_M_NLTVL1_W) std::printf("Before\nInitialization\n");
    return EXIT_SUCCESS;
}

// What main() prints after the solver returns.
void report_job(const Ready &R, bool verbose) {
    if (R.passthrough) return;
    if (verbose) {
        for (int k = 0; k < R.params.warps; k++) {
            if (R.method == FALDOI_M_TVL1_OCC)
                std::printf("Warping: %d, Iter: %d Error: %f\n", k, R.log.iters[k], R.log.err[k]);
            else if (is_nltv(R.method))
                std::printf("Warping: %d,Iter: %d Error: %f\n", k, R.log.iters[k], R.log.err[k]);
            else
                fprintf(stderr, "Warping: %d,Iter: %d Error: %f\n", k, R.log.iters[k], R.log.err[k]);
        }
    }
    if (R.method == FALDOI_M_TVL1 || R.method == FALDOI_M_TVL1_W) std::cout << "(tvl2OF) All tasks took " << R.secs << std::endl;
    if (R.method == FALDOI_M_TVCSAD || R.method == FALDOI_M_TVCSAD_W) std::printf("Exits current level\n");
}

// Stage 3 of a job: the output files.
void save_job(const Ready &R) {
    UsTimer ts(g_us_save);
    faldoi_host::write_image_float_split(R.flow_file, R.u, R.w, R.h, 2);
    if (R.method == FALDOI_M_TVL1_OCC && !R.passthrough) {
        std::vector<int> occ((size_t)R.w * R.h);
        for (size_t i = 0; i < occ.size(); i++) occ[i] = (int)R.chi[i];
        faldoi_host::write_png_gray8(R.occ_file, occ.data(), R.w, R.h);
    }
}

// Host-side preprocessing of one job (-host_preproc 1, and the stripe path, whose C ABI takes preprocessed frames)
struct HostFrames {
    std::vector<float> i0n, i1n, i_1n, lab;
};
HostFrames host_preprocess(const Ready &R) {
    HostFrames F;
    const size_t size = (size_t)R.w * R.h;
    F.i0n.resize(size), F.i1n.resize(size), F.i_1n.resize(size);
    if (is_nltv(R.method)) {
        F.lab.resize(3 * size);
        faldoi_host::image_to_lab(R.in.i0.px(), (int)size, F.lab.data());
    }
    faldoi_host::preprocess(R.in.i0.px(), R.in.i1.px(), R.in.i_1.px(), R.pd, R.w, R.h, F.i0n.data(), F.i1n.data(), F.i_1n.data());
    return F;
}

bool wants_stripes(const Ready &R, const Options &opt) {
    return opt.stripes && opt.devices.size() >= 2 && (R.method == FALDOI_M_TVL1 || R.method == FALDOI_M_TVL1_W) &&
           (long long)R.w * R.h >= 3840LL * 2160 && R.h >= 2 * (int)opt.devices.size();
}

// One pair, one call: everything between reading the files and saving the flow runs on the GPU (gray conversion,
// joint normalisation, Gaussian pre-smoothing, Lab and the minimisation).  Frames of 4K and more are cut into
// row stripes over the listed GPUs (TVL2).
int solve_single(Ready &R, const Options &opt) {
    if (R.passthrough) return FALDOI_OK;
    const auto t0 = std::chrono::system_clock::now();
    int rc;
    if (wants_stripes(R, opt)) {
        const HostFrames F = host_preprocess(R);
        faldoi_stripes *g = nullptr;
        rc = faldoi_stripes_create(&g, (int)opt.devices.size(), opt.devices.data(), R.w, R.h, R.method);
        if (rc == FALDOI_OK) rc = faldoi_stripes_upload(g, F.i0n.data(), F.i1n.data(), R.u);
        if (rc == FALDOI_OK) rc = faldoi_stripes_run(g, &R.params);
        if (rc == FALDOI_OK) rc = faldoi_stripes_download(g, R.u, &R.log);
        faldoi_stripes_destroy(g);
        if (rc == FALDOI_OK) fprintf(stderr, "row stripes: %d GPUs\n", (int)opt.devices.size());
    } else if (opt.host_preproc) {
        const HostFrames F = host_preprocess(R);
        rc = faldoi_global_solve(opt.device, &R.params, R.w, R.h, F.i0n.data(), F.i1n.data(), F.i_1n.data(), is_nltv(R.method) ? F.lab.data() : nullptr,
                                 R.u, R.method >= 8 ? R.chi : nullptr, &R.log);
    } else {
        rc = faldoi_global_solve_raw(opt.device, &R.params, R.w, R.h, R.pd, R.in.i0.px(), R.in.i1.px(), R.in.i_1.px(), R.u,
                                     R.method >= 8 ? R.chi : nullptr, &R.log);
    }
    R.secs = std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Sequence mode: loaders -> dispatcher (job order, main()'s checks) -> one worker per GPU (batches) -> writers.
// Every job in flight owns a page-locked staging slot from the moment its files are decoded until its results
// are on disk; loaders and writers are persistent pool threads (their scratch buffers are reused from job to job).
// ---------------------------------------------------------------------------------------------------------------
class ThreadPool {
  public:
    explicit ThreadPool(size_t n) {
        for (size_t i = 0; i < n; i++) th_.emplace_back([this] { loop(); });
    }
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    template <class F>
    auto submit(F f) -> std::future<decltype(f())> {
        auto task = std::make_shared<std::packaged_task<decltype(f())()>>(std::move(f));
        auto fut = task->get_future();
        {
            std::lock_guard<std::mutex> l(mu_);
            q_.push_back([task] { (*task)(); });
        }
        cv_.notify_one();
        return fut;
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return quit_ || !q_.empty(); });
                if (q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    std::vector<std::thread> th_;
    std::deque<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool quit_ = false;
};

class SlotPool {
  public:
    explicit SlotPool(size_t n) : slots_(n) {
        for (size_t i = 0; i < n; i++) free_.push_back((int)i);
    }
    int try_acquire() {
        std::lock_guard<std::mutex> l(mu_);
        if (free_.empty()) return -1;
        const int id = free_.back();
        free_.pop_back();
        return id;
    }
    int acquire() {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return !free_.empty(); });
        const int id = free_.back();
        free_.pop_back();
        return id;
    }
    void release(int id) {
        if (id < 0) return;
        {
            std::lock_guard<std::mutex> l(mu_);
            free_.push_back(id);
        }
        cv_.notify_one();
    }
    Slot *at(int id) { return &slots_[id]; }

  private:
    std::vector<Slot> slots_;
    std::vector<int> free_;
    std::mutex mu_;
    std::condition_variable cv_;
};

struct Batch {
    std::vector<std::unique_ptr<Ready>> jobs;
};

// One GPU: a host thread and a batched solver handle.
class DeviceWorker {
  public:
    DeviceWorker(int device, const Options &opt, ThreadPool &io, SlotPool &slots) : device_(device), opt_(opt), io_(io), slots_(slots) {
        th_ = std::thread([this] { loop(); });
    }
    ~DeviceWorker() { finish(); }
    // no more batches: the thread ends once its queue is empty and its results are on disk
    void finish() {
        {
            std::lock_guard<std::mutex> l(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }
    bool idle() {
        std::lock_guard<std::mutex> l(mu_);
        return !pending_ && !busy_;
    }
    // hands a batch over; blocks while the worker still has one waiting (queue of depth 1)
    void submit(std::unique_ptr<Batch> b) {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return !pending_; });
        pending_ = std::move(b);
        cv_.notify_all();
    }
    std::string error() {
        std::lock_guard<std::mutex> l(mu_);
        return err_;
    }
    int done() {
        std::lock_guard<std::mutex> l(mu_);
        return done_;
    }

  private:
    void loop() {
        for (;;) {
            std::unique_ptr<Batch> b;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return pending_ || quit_; });
                if (!pending_) break;
                b = std::move(pending_);
                busy_ = true;
            }
            cv_.notify_all();
            t_wait_ += since(t_mark_);
            std::string e;
            const int n = (int)b->jobs.size();
            try {
                e = run_batch(*b);
            } catch (const std::exception &ex) {
                e = ex.what();
            }
            for (auto &j : b->jobs)
                if (j) slots_.release(j->slot);  // (jobs that went to a writer were moved out of the batch)
            {
                std::lock_guard<std::mutex> l(mu_);
                if (!e.empty() && err_.empty()) err_ = e;
                if (e.empty()) done_ += n;
                busy_ = false;
            }
            cv_.notify_all();
            t_mark_ = std::chrono::steady_clock::now();
        }
        for (auto &w : writers_) collect(w);
        if (opt_.seq_stats)
            fprintf(stderr, "device %d: waiting for batches %.3f s, uploads %.3f s, solve %.3f s, downloads %.3f s, reports + writer hand-over %.3f s\n",
                    device_, t_wait_, t_up_, t_run_, t_down_, t_out_);
        if (solver_) faldoi_solver_destroy(solver_);
    }

    void collect(std::future<void> &w) {
        try {
            if (w.valid()) w.get();
        } catch (const std::exception &ex) {
            std::lock_guard<std::mutex> l(mu_);
            if (err_.empty()) err_ = ex.what();
        }
    }

    static double since(std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); }
    static std::string gpu_error(int rc) { return "GPU solver failed (" + std::to_string(rc) + "): " + faldoi_last_error(); }

    std::string run_batch(Batch &B) {
        const int n = (int)B.jobs.size();
        Ready &R0 = *B.jobs[0];
        if (!R0.passthrough) {
            // the handle: one per (size, method), with as many slots as a batch can have
            if (!solver_ || key_w_ != R0.w || key_h_ != R0.h || key_m_ != R0.method) {
                if (solver_) faldoi_solver_destroy(solver_);
                solver_ = nullptr;
                const int rc = faldoi_solver_create(&solver_, device_, R0.w, R0.h, R0.method, opt_.batch);
                if (rc != FALDOI_OK) return gpu_error(rc);
                key_w_ = R0.w, key_h_ = R0.h, key_m_ = R0.method;
            }
            const auto t0 = std::chrono::system_clock::now();
            auto tm = std::chrono::steady_clock::now();
            const bool occ = (R0.method == FALDOI_M_TVL1_OCC), nltv = is_nltv(R0.method);
            // uploads: asynchronous copies out of the jobs' page-locked staging slots, preprocessing on the device
            for (int k = 0; k < n; k++) {
                Ready &R = *B.jobs[k];
                int rc;
                if (opt_.host_preproc) {
                    const HostFrames F = host_preprocess(R);
                    rc = faldoi_solver_upload(solver_, k, F.i0n.data(), F.i1n.data(), F.i_1n.data(), nltv ? F.lab.data() : nullptr, R.u, occ ? R.chi : nullptr);
                    if (rc == FALDOI_OK) rc = faldoi_solver_sync(solver_);  // F goes out of scope
                } else {
                    rc = faldoi_solver_upload_raw(solver_, k, R.in.i0.px(), R.in.i1.px(), R.in.i_1.px(), R.pd, R.u, occ ? R.chi : nullptr);
                }
                if (rc != FALDOI_OK) return gpu_error(rc);
            }
            t_up_ += since(tm), tm = std::chrono::steady_clock::now();
            int rc = faldoi_solver_run(solver_, &R0.params, n);
            if (rc != FALDOI_OK) return gpu_error(rc);
            if (opt_.seq_stats) faldoi_solver_sync(solver_);
            t_run_ += since(tm), tm = std::chrono::steady_clock::now();
            for (int k = 0; k < n; k++) {
                Ready &R = *B.jobs[k];
                rc = faldoi_solver_download(solver_, k, R.u, occ ? R.chi : nullptr, &R.log);
                if (rc != FALDOI_OK) return gpu_error(rc);
            }
            t_down_ += since(tm);
            const double secs = std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
            for (int k = 0; k < n; k++) B.jobs[k]->secs = secs / n;
        }
        const auto to = std::chrono::steady_clock::now();
        {
            std::lock_guard<std::mutex> l(g_io_mu);
            for (int k = 0; k < n; k++) report_job(*B.jobs[k], opt_.verbose);
        }
        // results go to disk on the pool threads; a job's staging slot is free again once its files are written
        while (writers_.size() >= 4 * (size_t)opt_.batch) {
            collect(writers_.front());
            writers_.pop_front();
        }
        for (int k = 0; k < n; k++) {
            std::shared_ptr<Ready> job(std::move(B.jobs[k]));
            SlotPool *slots = &slots_;
            writers_.push_back(io_.submit([job, slots] {
                struct Release {
                    SlotPool *p;
                    int id;
                    ~Release() { p->release(id); }
                } rel{slots, job->slot};
                save_job(*job);
            }));
        }
        t_out_ += since(to);
        return std::string();
    }

    const int device_;
    const Options opt_;
    ThreadPool &io_;
    SlotPool &slots_;
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::unique_ptr<Batch> pending_;
    bool busy_ = false, quit_ = false;
    int done_ = 0;
    std::string err_;
    faldoi_solver *solver_ = nullptr;
    int key_w_ = 0, key_h_ = 0, key_m_ = -1;
    std::deque<std::future<void>> writers_;
    double t_wait_ = 0, t_up_ = 0, t_run_ = 0, t_down_ = 0, t_out_ = 0;  // -seq_stats 1
    std::chrono::steady_clock::time_point t_mark_ = std::chrono::steady_clock::now();
};

int run_sequence(const std::string &seq_file, const std::string &argv0, const Options &opt) {
    std::ifstream jobs(seq_file);
    if (!jobs) {
        fprintf(stderr, "ERROR: cannot open job list '%s'\n", seq_file.c_str());
        return EXIT_FAILURE;
    }
    std::string line;
    std::vector<std::vector<std::string>> joblist;
    while (std::getline(jobs, line)) {
        std::vector<std::string> job{argv0};
        std::string tok;
        for (std::istringstream ls(line); ls >> tok;) job.push_back(tok);
        if (job.size() == 1) continue;
        if (job.size() != 4 && job.size() != 6) {
            fprintf(stderr, "ERROR: job %d of '%s' needs 3 or 5 file names\n", (int)joblist.size() + 1, seq_file.c_str());
            return EXIT_FAILURE;
        }
        joblist.push_back(job);
    }
    const size_t njobs = joblist.size(), ndev = opt.devices.size();
    // per GPU: one batch on the device, one queued or being assembled, one on its way to disk
    const size_t nslots = std::max<size_t>(3 * (size_t)opt.batch * ndev, 8);
    const bool pinned = g_have_gpu && (opt.pinned == 1 || (opt.pinned < 0 && njobs >= 20 * nslots));
    PinnedArenas arenas(4, pinned);
    g_arenas = &arenas;
    SlotPool slots(nslots);
    const size_t window = std::max<size_t>(2 * (size_t)opt.batch * ndev, 4);
    ThreadPool io(std::max(4u, std::min(64u, std::thread::hardware_concurrency())));
    int rc = EXIT_SUCCESS;
    int done = 0;
    std::string err;
    const auto t_seq0 = std::chrono::steady_clock::now();
    {
        std::vector<std::unique_ptr<DeviceWorker>> workers;
        for (int d : opt.devices) workers.emplace_back(new DeviceWorker(d, opt, io, slots));

        struct InFlight {
            std::future<Loaded> fut;
            int slot;
        };
        std::deque<InFlight> loading;
        size_t next_load = 0, rr = 0;
        std::unique_ptr<Batch> cur;
        auto start_load = [&](int slot) {
            const size_t k = next_load++;
            Slot *sp = slots.at(slot);
            loading.push_back(InFlight{io.submit([&joblist, &opt, k, sp] { return load_inputs(joblist[k], opt.val_method, false, sp); }), slot});
        };
        auto refill = [&] {  // never blocks: decode ahead only into staging slots that are free
            while (next_load < njobs && loading.size() < window) {
                const int slot = slots.try_acquire();
                if (slot < 0) break;
                start_load(slot);
            }
        };
        auto flush = [&] {
            if (!cur || cur->jobs.empty()) return;
            // an idle GPU if there is one, else round robin (submit blocks while that worker's queue slot is taken)
            size_t pick = rr;
            for (size_t i = 0; i < workers.size(); i++)
                if (workers[(rr + i) % workers.size()]->idle()) {
                    pick = (rr + i) % workers.size();
                    break;
                }
            rr = (pick + 1) % workers.size();
            workers[pick]->submit(std::move(cur));
            cur.reset();
        };
        auto worker_failed = [&] {
            for (auto &w : workers)
                if (!w->error().empty()) return true;
            return false;
        };
        for (size_t k = 0; k < njobs && rc == EXIT_SUCCESS; k++) {
            refill();
            if (loading.empty()) {
                // every slot is taken by jobs further down the pipeline: hand over what has been assembled (so that
                // slots are certain to come back) and wait for one
                flush();
                start_load(slots.acquire());
            }
            Loaded in;
            {
                UsTimer tw(g_us_wait_load);
                in = loading.front().fut.get();
            }
            const int slot = loading.front().slot;
            loading.pop_front();
            std::unique_ptr<Ready> R(new Ready());
            R->slot = slot;
            try {
                UsTimer tp(g_us_prepare);
                std::lock_guard<std::mutex> l(g_io_mu);
                const int r = prepare_job(joblist[k], opt, std::move(in), *R);
                if (r != EXIT_SUCCESS) rc = r;
            } catch (const std::exception &e) {
                fprintf(stderr, "ERROR: %s\n", e.what());
                rc = EXIT_FAILURE;
            }
            if (rc != EXIT_SUCCESS || worker_failed()) {
                slots.release(slot);
                break;
            }
            const bool fits = cur && !cur->jobs.empty() && cur->jobs[0]->w == R->w && cur->jobs[0]->h == R->h && cur->jobs[0]->pd == R->pd &&
                              cur->jobs[0]->method == R->method && cur->jobs[0]->passthrough == R->passthrough;
            if (!fits) flush();
            if (!cur) cur.reset(new Batch());
            cur->jobs.push_back(std::move(R));
            if ((int)cur->jobs.size() == opt.batch) flush();
        }
        flush();  // the jobs before a failing one are completed
        for (auto &f : loading) {
            if (f.fut.valid()) f.fut.wait();
            slots.release(f.slot);
        }
        for (auto &w : workers) w->finish();
        for (auto &w : workers) {
            done += w->done();
            if (err.empty()) err = w->error();
        }
    }
    if (opt.seq_stats) {
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_seq0).count();
        fprintf(stderr, "pipeline: %d pairs in %.3f s = %.1f pairs/s, files in to files out (handle creation included, process and CUDA start-up not)\n",
                done, secs, done / secs);
    }
    if (opt.seq_stats)
        fprintf(stderr, "host: dispatcher waited for decoded jobs %.3f s, checks %.3f s; summed over the I/O threads: decode %.3f s, page-locking %.3f s, "
                        "writing %.3f s\n", g_us_wait_load / 1e6, g_us_prepare / 1e6, g_us_decode / 1e6, g_us_pin / 1e6, g_us_save / 1e6);
    if (!err.empty()) {
        fprintf(stderr, "ERROR: %s\n", err.c_str());
        return EXIT_FAILURE;
    }
    if (rc != EXIT_SUCCESS) return rc;
    fprintf(stderr, "sequence: %d pairs done\n", done);
    return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char *argv[]) {
    // Frames and flows are multi-megabyte arrays allocated and freed at a high rate by many threads: keep them on
    // the heap instead of one mmap/munmap pair (and its TLB shoot-down across all threads) per array.
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    print_today();
    std::vector<std::string> args(argv, argv + argc);
    const std::string warps_val = pick_option(args, "w", "5");
    const std::string method_val = pick_option(args, "m", "0");
    const std::string file_params = pick_option(args, "p", "");
    const std::string global_iters = pick_option(args, "glb_iters", "400");
    const std::string verbose_str = pick_option(args, "verbose", "0");
    const std::string device_str = pick_option(args, "device", "");
    const std::string devices_str = pick_option(args, "devices", "");
    const std::string batch_str = pick_option(args, "batch", "16");
    const bool host_preproc = pick_option(args, "host_preproc", "0") == "1";
    const bool stripes = pick_option(args, "stripes", "1") != "0";
    const bool seq_stats = pick_option(args, "seq_stats", "0") == "1";
    const std::string pinned_str = pick_option(args, "pinned", "-1");
    const std::string seq_file = pick_option(args, "seq", "");

    if (seq_file.empty() && args.size() != 6 && args.size() != 4) {
        usage(args.size());
        return EXIT_FAILURE;
    }

    Options opt;
    try {
        opt.val_method = std::stoi(method_val);
        opt.nwarps = std::stoi(warps_val);
        opt.glb_it = std::stoi(global_iters);
        opt.device = device_str.empty() ? 0 : std::stoi(device_str);
        opt.batch = std::max(1, std::stoi(batch_str));
        opt.pinned = std::stoi(pinned_str);
        if (verbose_str != "0" && verbose_str != "1") throw std::invalid_argument("-verbose takes 0 or 1");
        opt.verbose = (verbose_str == "1");
        // GPUs: -devices list | "all"; else -device d; else device 0.  A single call without either option may
        // use every visible GPU (row stripes of a >= 4K frame).
        const int visible = faldoi_device_count();
        g_have_gpu = visible > 0;
        if (devices_str == "all" || (devices_str.empty() && device_str.empty() && seq_file.empty())) {
            for (int d = 0; d < visible; d++) opt.devices.push_back(d);
        } else if (!devices_str.empty()) {
            std::string tok;
            for (std::istringstream ls(devices_str); std::getline(ls, tok, ',');) opt.devices.push_back(std::stoi(tok));
        }
        if (opt.devices.empty()) opt.devices.push_back(opt.device);
        for (int d : opt.devices)
            if (d < 0 || (visible > 0 && d >= visible)) throw std::invalid_argument("no such CUDA device: " + std::to_string(d));
        if (!devices_str.empty() && devices_str != "all" && device_str.empty()) opt.device = opt.devices[0];
    } catch (const std::exception &e) {
        fprintf(stderr, "ERROR: bad option value (%s)\n", e.what());
        return EXIT_FAILURE;
    }
    opt.host_preproc = host_preproc;
    opt.stripes = stripes;
    opt.seq_stats = seq_stats;
    opt.file_params = file_params;

    int rc = EXIT_SUCCESS;
    if (seq_file.empty()) {
        Ready R;
        try {
            rc = prepare_job(args, opt, load_inputs(args, opt.val_method, true), R);
            if (rc == EXIT_SUCCESS) {
                const int g = solve_single(R, opt);
                if (g != FALDOI_OK) {
                    fprintf(stderr, "ERROR: GPU solver failed (%d): %s\n", g, faldoi_last_error());
                    return EXIT_FAILURE;
                }
                report_job(R, opt.verbose);
                save_job(R);
            }
        } catch (const std::exception &e) {
            fprintf(stderr, "ERROR: %s\n", e.what());
            return EXIT_FAILURE;
        }
    } else {
        rc = run_sequence(seq_file, args[0], opt);
    }
    if (rc != EXIT_SUCCESS) return rc;
    print_today();
    return EXIT_SUCCESS;
}
