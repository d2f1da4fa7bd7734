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
//   -host_preproc 1  run main()'s preprocessing on the host instead of the GPU
//   -seq jobs.txt    many pairs in one process (one job per line: the positional arguments of a normal call).
//                    Files are read and decoded by a pool of host threads, jobs of equal size fill the slots of a
//                    batched solver handle (-batch n, default 16) through pinned staging buffers, results are
//                    written by background threads -- SURVEY 8f rows 3 and 4; what the reference's scripts do with
//                    one process per pair (scripts_python/faldoi_sift.py:314-318)
// There is no CPU fallback: without a usable GPU the program reports the error and fails.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
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
    int val_method = 0, nwarps = 5, glb_it = 400, device = 0, batch = 16;
    bool verbose = false, host_preproc = false, stripes = true;
    std::vector<int> devices;
    std::string file_params;
};

// Stage 1 of a job (host I/O, no messages): ims.txt + the frames, the flow and the occlusion
// mask.  A single call decodes its files concurrently; a sequence decodes many jobs at once
// instead.  Errors are kept for the solve stage to report in job order.
struct Loaded {
    Image i_1, i0, i1, flow, occ;
    int num_files = 0;
    std::exception_ptr err;
};

Loaded load_inputs(const std::vector<std::string> &args, int val_method, bool parallel) {
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
        auto rd = [](std::string f) { return faldoi_host::read_image_split(f); };
        const auto policy = parallel ? std::launch::async : std::launch::deferred;
        std::future<Image> f0 = std::async(policy, rd, filename_i0);
        std::future<Image> f1 = std::async(policy, rd, filename_i1);
        std::future<Image> ff = std::async(policy, rd, args[2]);
        std::future<Image> fo;
        if (val_method >= 8) fo = std::async(policy, rd, args.size() == 6 ? args[4] : std::string());
        std::exception_ptr first;
        auto take = [&](std::future<Image> &f, Image &dst) {
            try {
                dst = f.get();
            } catch (...) {
                if (!first) first = std::current_exception();
            }
        };
        const bool third_is_i1 = (third == filename_i1);  // the same file: decode it once
        try {  // same order as the reference reads them, so the same file is blamed first
            if (!third_is_i1) L.i_1 = rd(third);
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
    std::vector<float> u, chi;
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
    R.u = std::move(R.in.flow.data);  // u1 | u2
    if (val_method >= 8 && !R.passthrough) R.chi.assign(occ.data.begin(), occ.data.begin() + size);
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
    faldoi_host::write_image_float_split(R.flow_file, R.u.data(), R.w, R.h, 2);
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
        faldoi_host::image_to_lab(R.in.i0.data.data(), (int)size, F.lab.data());
    }
    faldoi_host::preprocess(R.in.i0.data.data(), R.in.i1.data.data(), R.in.i_1.data.data(), R.pd, R.w, R.h, F.i0n.data(), F.i1n.data(),
                            F.i_1n.data());
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
        if (rc == FALDOI_OK) rc = faldoi_stripes_upload(g, F.i0n.data(), F.i1n.data(), R.u.data());
        if (rc == FALDOI_OK) rc = faldoi_stripes_run(g, &R.params);
        if (rc == FALDOI_OK) rc = faldoi_stripes_download(g, R.u.data(), &R.log);
        faldoi_stripes_destroy(g);
        if (rc == FALDOI_OK) fprintf(stderr, "row stripes: %d GPUs\n", (int)opt.devices.size());
    } else if (opt.host_preproc) {
        const HostFrames F = host_preprocess(R);
        rc = faldoi_global_solve(opt.device, &R.params, R.w, R.h, F.i0n.data(), F.i1n.data(), F.i_1n.data(), is_nltv(R.method) ? F.lab.data() : nullptr,
                                 R.u.data(), R.method >= 8 ? R.chi.data() : nullptr, &R.log);
    } else {
        rc = faldoi_global_solve_raw(opt.device, &R.params, R.w, R.h, R.pd, R.in.i0.data.data(), R.in.i1.data.data(), R.in.i_1.data.data(),
                                     R.u.data(), R.method >= 8 ? R.chi.data() : nullptr, &R.log);
    }
    R.secs = std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Sequence mode: loaders -> dispatcher (job order, main()'s checks) -> one worker per GPU (batches) -> writers
// ---------------------------------------------------------------------------------------------------------------
struct Batch {
    std::vector<std::unique_ptr<Ready>> jobs;
};

// One GPU: a host thread, a batched solver handle and pinned staging for its slots.
class DeviceWorker {
  public:
    DeviceWorker(int device, const Options &opt) : device_(device), opt_(opt) { th_ = std::thread([this] { loop(); }); }
    ~DeviceWorker() { finish(); }
    // no more batches: the thread ends once its queue is empty and its background writers are done
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
    void drain() {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return !pending_ && !busy_; });
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
            std::string e;
            const int n = (int)b->jobs.size();
            try {
                e = run_batch(*b);
            } catch (const std::exception &ex) {
                e = ex.what();
            }
            {
                std::lock_guard<std::mutex> l(mu_);
                if (!e.empty() && err_.empty()) err_ = e;
                if (e.empty()) done_ += n;
                busy_ = false;
            }
            cv_.notify_all();
        }
        for (auto &w : writers_)
            if (w.valid()) {
                try {
                    w.get();
                } catch (const std::exception &ex) {
                    std::lock_guard<std::mutex> l(mu_);
                    if (err_.empty()) err_ = ex.what();
                }
            }
        if (solver_) faldoi_solver_destroy(solver_);
        faldoi_pinned_free(pin_);
    }

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
            // pinned staging, per slot: three frames (pd planes each; at least 2 planes so that host-preprocessed
            // gray + Lab fit), the flow (in and out) and chi
            const size_t size = (size_t)R0.w * R0.h, pd = R0.pd, fpl = std::max<size_t>(pd, 2);
            const size_t per_slot = 3 * fpl * size + 2 * size + size;
            if (per_slot * opt_.batch > pin_floats_) {
                faldoi_pinned_free(pin_);
                pin_ = (float *)faldoi_pinned_alloc(per_slot * opt_.batch * sizeof(float));
                if (!pin_) return std::string("pinned staging allocation failed: ") + faldoi_last_error();
                pin_floats_ = per_slot * opt_.batch;
            }
            const auto t0 = std::chrono::system_clock::now();
            const bool occ = (R0.method == FALDOI_M_TVL1_OCC), nltv = is_nltv(R0.method);
            for (int k = 0; k < n; k++) {
                Ready &R = *B.jobs[k];
                float *p = pin_ + per_slot * k, *pu = p + 3 * fpl * size, *pc = pu + 2 * size;
                memcpy(pu, R.u.data(), 2 * size * sizeof(float));
                if (occ) memcpy(pc, R.chi.data(), size * sizeof(float));
                int rc;
                if (opt_.host_preproc) {
                    const HostFrames F = host_preprocess(R);
                    memcpy(p, F.i0n.data(), size * sizeof(float));
                    memcpy(p + size, F.i1n.data(), size * sizeof(float));
                    memcpy(p + 2 * size, F.i_1n.data(), size * sizeof(float));
                    if (nltv) memcpy(p + 3 * size, F.lab.data(), 3 * size * sizeof(float));
                    rc = faldoi_solver_upload(solver_, k, p, p + size, p + 2 * size, nltv ? p + 3 * size : nullptr, pu, occ ? pc : nullptr);
                } else {
                    memcpy(p, R.in.i0.data.data(), pd * size * sizeof(float));
                    memcpy(p + pd * size, R.in.i1.data.data(), pd * size * sizeof(float));
                    memcpy(p + 2 * pd * size, R.in.i_1.data.data(), pd * size * sizeof(float));
                    rc = faldoi_solver_upload_raw(solver_, k, p, p + pd * size, p + 2 * pd * size, (int)pd, pu, occ ? pc : nullptr);
                }
                if (rc != FALDOI_OK) return gpu_error(rc);
            }
            int rc = faldoi_solver_run(solver_, &R0.params, n);
            if (rc != FALDOI_OK) return gpu_error(rc);
            for (int k = 0; k < n; k++) {
                Ready &R = *B.jobs[k];
                float *pu = pin_ + per_slot * k + 3 * fpl * size, *pc = pu + 2 * size;
                rc = faldoi_solver_download(solver_, k, pu, occ ? pc : nullptr, &R.log);
                if (rc != FALDOI_OK) return gpu_error(rc);
                memcpy(R.u.data(), pu, 2 * size * sizeof(float));
                if (occ) memcpy(R.chi.data(), pc, size * sizeof(float));
            }
            const double secs = std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
            for (int k = 0; k < n; k++) B.jobs[k]->secs = secs / n;
        }
        {
            std::lock_guard<std::mutex> l(g_io_mu);
            for (int k = 0; k < n; k++) report_job(*B.jobs[k], opt_.verbose);
        }
        // results go to disk on background threads (at most two batches' files in flight per GPU)
        while (writers_.size() >= 2 * (size_t)opt_.batch) {
            writers_.front().get();
            writers_.pop_front();
        }
        for (int k = 0; k < n; k++) {
            std::shared_ptr<Ready> job(std::move(B.jobs[k]));
            job->in = Loaded();  // the frames are no longer needed
            writers_.push_back(std::async(std::launch::async, [job] { save_job(*job); }));
        }
        return std::string();
    }

    const int device_;
    const Options opt_;
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::unique_ptr<Batch> pending_;
    bool busy_ = false, quit_ = false;
    int done_ = 0;
    std::string err_;
    faldoi_solver *solver_ = nullptr;
    int key_w_ = 0, key_h_ = 0, key_m_ = -1;
    float *pin_ = nullptr;
    size_t pin_floats_ = 0;
    std::deque<std::future<void>> writers_;
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
    const size_t njobs = joblist.size();
    std::vector<std::unique_ptr<DeviceWorker>> workers;
    for (int d : opt.devices) workers.emplace_back(new DeviceWorker(d, opt));

    // loaders: a window of jobs is being read and decoded ahead of the dispatcher, on a bounded number of threads
    const size_t window = std::max<size_t>(2 * (size_t)opt.batch * workers.size(), 8);
    const size_t nthreads = std::max(4u, std::min(48u, std::thread::hardware_concurrency()));
    std::deque<std::future<Loaded>> loading;
    size_t next_load = 0, pool_busy = 0;
    std::mutex pool_mu;
    std::condition_variable pool_cv;
    auto start_load = [&](size_t k) {
        {
            std::unique_lock<std::mutex> l(pool_mu);
            pool_cv.wait(l, [&] { return pool_busy < nthreads; });
            pool_busy++;
        }
        return std::async(std::launch::async, [&, k] {
            Loaded L = load_inputs(joblist[k], opt.val_method, false);
            {
                std::lock_guard<std::mutex> l(pool_mu);
                pool_busy--;
            }
            pool_cv.notify_one();
            return L;
        });
    };
    auto refill = [&] {
        while (next_load < njobs && loading.size() < window) loading.push_back(start_load(next_load++));
    };

    int rc = EXIT_SUCCESS;
    size_t rr = 0;
    std::unique_ptr<Batch> cur;
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
    refill();
    for (size_t k = 0; k < njobs && rc == EXIT_SUCCESS; k++) {
        Loaded in = loading.front().get();
        loading.pop_front();
        refill();
        std::unique_ptr<Ready> R(new Ready());
        try {
            std::lock_guard<std::mutex> l(g_io_mu);
            const int r = prepare_job(joblist[k], opt, std::move(in), *R);
            if (r != EXIT_SUCCESS) rc = r;
        } catch (const std::exception &e) {
            fprintf(stderr, "ERROR: %s\n", e.what());
            rc = EXIT_FAILURE;
        }
        if (rc != EXIT_SUCCESS || worker_failed()) break;
        const bool fits = cur && !cur->jobs.empty() && cur->jobs[0]->w == R->w && cur->jobs[0]->h == R->h && cur->jobs[0]->pd == R->pd &&
                          cur->jobs[0]->method == R->method && cur->jobs[0]->passthrough == R->passthrough;
        if (!fits) flush();
        if (!cur) cur.reset(new Batch());
        cur->jobs.push_back(std::move(R));
        if ((int)cur->jobs.size() == opt.batch) flush();
    }
    flush();  // the jobs before a failing one are completed
    for (auto &f : loading)
        if (f.valid()) f.wait();
    for (auto &w : workers) w->finish();
    int done = 0;
    std::string err;
    for (auto &w : workers) {
        done += w->done();
        if (err.empty()) err = w->error();
    }
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
        if (verbose_str != "0" && verbose_str != "1") throw std::invalid_argument("-verbose takes 0 or 1");
        opt.verbose = (verbose_str == "1");
        // GPUs: -devices list | "all"; else -device d; else device 0.  A single call without either option may
        // use every visible GPU (row stripes of a >= 4K frame).
        const int visible = faldoi_device_count();
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
