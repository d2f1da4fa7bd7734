// global_faldoi -- drop-in for the reference executable (src/global_faldoi.cpp:1846-2213):
//
//   global_faldoi ims.txt in_flow.flo out.flo [occ_in.png occ_out.png]
//                 [-m method] [-w warps] [-p params.txt] [-glb_iters n] [-verbose 0|1]
//
// Same argv contract, file formats, method ids, parameter defaults and messages; the
// minimisation itself (tvl2OF / nltvl1_PD / tvcsad_PD / nltvcsad_PD /
// guided_tvl2coupled_occ) runs on a B200 through the C ABI in include/faldoi_gpu.h.
// Extra options (unknown to the reference's scripts): -device d (CUDA device, default 0),
// -host_preproc 1 (run main()'s preprocessing on the host instead of the GPU), -seq jobs.txt
// (many pairs in one process; files of job k+1 are read and decoded and the results of job k-1
// are written by host threads while job k is on the GPU -- SURVEY 8f rows 3 and 4).
// There is no CPU fallback: without a usable GPU the program reports the error and fails.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
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

}  // namespace

struct Options {
    int val_method, nwarps, glb_it, device;
    bool verbose, host_preproc;
    std::string file_params;
};

// Stage 1 of a job (host I/O, no messages): ims.txt + the frames, the flow and the occlusion
// mask, decoded concurrently.  Errors are kept for the solve stage to report in job order.
struct Loaded {
    Image i_1, i0, i1, flow, occ;
    int num_files = 0;
    std::exception_ptr err;
};

static Loaded load_inputs(const std::vector<std::string> &args, int val_method) {
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
        std::future<Image> f0 = std::async(std::launch::async, rd, filename_i0);
        std::future<Image> f1 = std::async(std::launch::async, rd, filename_i1);
        std::future<Image> ff = std::async(std::launch::async, rd, args[2]);
        std::future<Image> fo;
        if (val_method >= 8) fo = std::async(std::launch::async, rd, args.size() == 6 ? args[4] : std::string());
        std::exception_ptr first;
        auto take = [&](std::future<Image> &f, Image &dst) {
            try {
                dst = f.get();
            } catch (...) {
                if (!first) first = std::current_exception();
            }
        };
        try {  // same order as the reference reads them, so the same file is blamed first
            L.i_1 = rd(third);
        } catch (...) {
            first = std::current_exception();
        }
        take(f0, L.i0);
        take(f1, L.i1);
        take(ff, L.flow);
        if (val_method >= 8) take(fo, L.occ);
        L.err = first;
    } catch (...) {
        L.err = std::current_exception();
    }
    return L;
}

// Stage 3 of a job: the output files.
struct Outputs {
    std::string flow_file, occ_file;
    std::vector<float> u;
    std::vector<int> occ;
    int w = 0, h = 0;
};
static void save_outputs(const Outputs &o) {
    faldoi_host::write_image_float_split(o.flow_file, o.u.data(), o.w, o.h, 2);
    if (!o.occ_file.empty()) faldoi_host::write_png_gray8(o.occ_file, o.occ.data(), o.w, o.h);
}

// one invocation of the reference executable: positional = {argv0, ims.txt, in.flo, out.flo[, occ_in, occ_out]}.
// `writer` (sequence mode): the files are written by a host thread while the next job is solved.
static int run_pair(const std::vector<std::string> &args, const Options &opt, Loaded &in, std::future<void> *writer) {
    int val_method = opt.val_method;
    const int nwarps = opt.nwarps, glb_it = opt.glb_it, device = opt.device;
    const bool verbose = opt.verbose, host_preproc = opt.host_preproc;
    const std::string &file_params = opt.file_params;
    const std::string &outfile = args[3];
    std::string occ_output;
    if (args.size() == 6) occ_output = args[5];
    const int num_files = in.num_files;

    try {
        if (in.err) std::rethrow_exception(in.err);
        const Image &i_1 = in.i_1, &i0 = in.i0, &i1 = in.i1, &flow = in.flow, &occ = in.occ;

        auto same = [](const Image &a, const Image &b) { return a.w == b.w && a.h == b.h && a.pd == b.pd; };
        if (num_files == 3) {
            if (!same(i0, i1) || !same(i0, i_1) || !same(i1, i_1))
                return fprintf(stderr, "ERROR: input images and flow size mismatch\n");
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
        if (val_method < 0 || val_method > 8) val_method = FALDOI_M_TVL1;  // the reference's dispatch falls through to nothing; we solve TVL2

        const int w = i0.w, h = i0.h, pd = i0.pd;
        const size_t size = (size_t)w * h;
        if (val_method == FALDOI_M_TVL1_OCC && (occ.w != w || occ.h != h))
            return fprintf(stderr, "ERROR: input images and flow size mismatch\n");

        faldoi_params params;
        if (faldoi_params_from_file(file_params.c_str(), val_method, glb_it, &params) != FALDOI_OK) {
            fprintf(stderr, "ERROR: %s\n", faldoi_last_error());
            return EXIT_FAILURE;
        }
        params.warps = nwarps;
        if (verbose)
            std::cerr << "Parameters: \n lambda: " << params.lambda << ", theta: " << params.theta << ", beta: " << params.beta
                      << ", alpha: " << params.alpha << ", \n tau_u: " << params.tau_u << ", tau_eta: " << params.tau_eta
                      << ", tau_chi: " << params.tau_chi << ", mu: " << params.mu << "\n";

        const bool nltv = (val_method == FALDOI_M_NLTVL1 || val_method == FALDOI_M_NLTVL1_W ||
                           val_method == FALDOI_M_NLTVCSAD || val_method == FALDOI_M_NLTVCSAD_W);
        if (nltv) {
            std::printf("W:%d H:%d Pd:%d\n", w, h, pd);
            if (pd < 3) {
                fprintf(stderr, "ERROR: the NLTV models need a colour (3-channel) first frame\n");
                return EXIT_FAILURE;
            }
        }
        if (w < 5 || h < 5) {  // gaussian() aborts with "sigma too large" on such frames
            fprintf(stderr, "GaussianSmooth: sigma too large\n");
            return EXIT_FAILURE;
        }

        std::vector<float> u(std::move(in.flow.data));  // u1 | u2
        std::vector<float> chi;
        if (val_method >= 8) chi.assign(occ.data.begin(), occ.data.begin() + size);

        if (nltv && (val_method == FALDOI_M_NLTVL1 || val_method == FALDOI_M_NLTVL1_W)) std::printf("Before\nInitialization\n");

        // Everything between reading the files and saving the flow runs on the GPU: gray conversion,
        // joint normalisation, Gaussian pre-smoothing, Lab (NLTV) and the minimisation itself.
        // -host_preproc 1 keeps main()'s preprocessing on the host (same bits for gray/normalise/smooth).
        const auto t0 = std::chrono::system_clock::now();
        faldoi_log log{};
        int rc;
        if (host_preproc) {
            std::vector<float> lab, i0n(size), i1n(size), i_1n(size);
            if (nltv) {
                lab.resize(3 * size);
                faldoi_host::image_to_lab(i0.data.data(), (int)size, lab.data());
            }
            faldoi_host::preprocess(i0.data.data(), i1.data.data(), i_1.data.data(), pd, w, h, i0n.data(), i1n.data(), i_1n.data());
            rc = faldoi_global_solve(device, &params, w, h, i0n.data(), i1n.data(), i_1n.data(), nltv ? lab.data() : nullptr, u.data(),
                                     val_method >= 8 ? chi.data() : nullptr, &log);
        } else {
            rc = faldoi_global_solve_raw(device, &params, w, h, pd, i0.data.data(), i1.data.data(), i_1.data.data(), u.data(),
                                         val_method >= 8 ? chi.data() : nullptr, &log);
        }
        if (rc != FALDOI_OK) {
            fprintf(stderr, "ERROR: GPU solver failed (%d): %s\n", rc, faldoi_last_error());
            return EXIT_FAILURE;
        }
        const std::chrono::duration<double> secs = std::chrono::system_clock::now() - t0;

        if (verbose) {
            for (int k = 0; k < params.warps; k++) {
                if (val_method == FALDOI_M_TVL1_OCC)
                    std::printf("Warping: %d, Iter: %d Error: %f\n", k, log.iters[k], log.err[k]);
                else if (nltv)
                    std::printf("Warping: %d,Iter: %d Error: %f\n", k, log.iters[k], log.err[k]);
                else
                    fprintf(stderr, "Warping: %d,Iter: %d Error: %f\n", k, log.iters[k], log.err[k]);
            }
        }
        if (val_method == FALDOI_M_TVL1 || val_method == FALDOI_M_TVL1_W) std::cout << "(tvl2OF) All tasks took " << secs.count() << std::endl;
        if (val_method == FALDOI_M_TVCSAD || val_method == FALDOI_M_TVCSAD_W) std::printf("Exits current level\n");

        Outputs out;
        out.flow_file = outfile;
        out.u = std::move(u);
        out.w = w, out.h = h;
        if (val_method == FALDOI_M_TVL1_OCC) {
            out.occ_file = occ_output;
            out.occ.resize(size);
            for (size_t i = 0; i < size; i++) out.occ[i] = (int)chi[i];
        }
        if (writer) {
            if (writer->valid()) writer->get();  // at most one job's files in flight; its errors surface here
            *writer = std::async(std::launch::async, [o = std::move(out)]() { save_outputs(o); });
        } else {
            save_outputs(out);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "ERROR: %s\n", e.what());
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

int main(int argc, char *argv[]) {
    print_today();
    std::vector<std::string> args(argv, argv + argc);
    const std::string warps_val = pick_option(args, "w", "5");
    const std::string method_val = pick_option(args, "m", "0");
    const std::string file_params = pick_option(args, "p", "");
    const std::string global_iters = pick_option(args, "glb_iters", "400");
    const std::string verbose_str = pick_option(args, "verbose", "0");
    const std::string device_str = pick_option(args, "device", "0");
    const bool host_preproc = pick_option(args, "host_preproc", "0") == "1";
    // Sequence mode (not in the reference): -seq jobs.txt, one job per line with the positional
    // arguments of a normal call ("ims.txt in.flo out.flo [occ_in.png occ_out.png]").  All jobs run
    // in this process with the options given on the command line, so CUDA start-up and the HBM
    // allocations are paid once per sequence instead of once per pair.
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
        opt.device = std::stoi(device_str);
        if (verbose_str != "0" && verbose_str != "1") throw std::invalid_argument("-verbose takes 0 or 1");
        opt.verbose = (verbose_str == "1");
    } catch (const std::exception &e) {
        fprintf(stderr, "ERROR: bad option value (%s)\n", e.what());
        return EXIT_FAILURE;
    }
    opt.host_preproc = host_preproc;
    opt.file_params = file_params;

    int rc = EXIT_SUCCESS;
    if (seq_file.empty()) {
        Loaded in = load_inputs(args, opt.val_method);
        rc = run_pair(args, opt, in, nullptr);
    } else {
        std::ifstream jobs(seq_file);
        if (!jobs) {
            fprintf(stderr, "ERROR: cannot open job list '%s'\n", seq_file.c_str());
            return EXIT_FAILURE;
        }
        std::string line;
        std::vector<std::vector<std::string>> joblist;
        while (std::getline(jobs, line)) {
            std::vector<std::string> job{args[0]};
            std::string tok;
            for (std::istringstream ls(line); ls >> tok;) job.push_back(tok);
            if (job.size() == 1) continue;
            if (job.size() != 4 && job.size() != 6) {
                fprintf(stderr, "ERROR: job %d of '%s' needs 3 or 5 file names\n", (int)joblist.size() + 1, seq_file.c_str());
                return EXIT_FAILURE;
            }
            joblist.push_back(job);
        }
        // three-stage pipeline: load k+1 | solve k (this thread, GPU) | save k-1
        int njobs = 0;
        std::future<void> writer;
        std::future<Loaded> next;
        if (!joblist.empty()) next = std::async(std::launch::async, load_inputs, joblist[0], opt.val_method);
        for (size_t k = 0; k < joblist.size(); k++) {
            Loaded in = next.get();
            if (k + 1 < joblist.size()) next = std::async(std::launch::async, load_inputs, joblist[k + 1], opt.val_method);
            const int r = run_pair(joblist[k], opt, in, &writer);
            if (r != EXIT_SUCCESS) {
                if (next.valid()) next.wait();
                if (writer.valid()) writer.wait();
                return r;
            }
            njobs++;
        }
        try {
            if (writer.valid()) writer.get();
        } catch (const std::exception &e) {
            fprintf(stderr, "ERROR: %s\n", e.what());
            return EXIT_FAILURE;
        }
        fprintf(stderr, "sequence: %d pairs done\n", njobs);
    }
    if (rc != EXIT_SUCCESS) return rc;
    print_today();
    return EXIT_SUCCESS;
}
