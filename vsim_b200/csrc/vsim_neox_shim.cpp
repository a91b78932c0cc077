// vsim_neox_shim.cpp -- the reference-side binding of the device-resident GPT-NeoX graph:
// what a maintainer adds to vsim.cpp so that main_gptneox (vsim.cpp:749-910) evaluates tokens through b200neox_eval
// instead of building and running the ggml graph (gptneox_eval, vsim.cpp:470-747).  Three lines of vsim.cpp change
// ("device-resident graph"; applied by sed in oracle/Makefile for the integration tests):
//
//     bool b200_dev_open(const char *model_path, int n_ctx);                       // declaration, above gptneox_eval
//     bool b200_dev_eval(int, int, const int *, int, std::vector<float> &, size_t &, int);
//     ...
//     if (b200_dev_eval(n_threads, n_past, embd_inp.data(), (int)embd_inp.size(), embd_w, mem_per_token, model.hparams.n_vocab)) return true;
//                                                                                   // first statement of gptneox_eval
//     b200_dev_open(params.model.c_str(), model.hparams.n_ctx);                     // after init_xmax() in main_gptneox
//
// B200NEOX=0 leaves the program on its own graph (then only the Q4 mul_mats run on the GPU, through the IMAX hook).
// B200Q4_EXACT=1 makes the device graph follow the reference's arithmetic exactly: same logits, same tokens.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/b200q4.h"
#include "../../include/b200neox.h"

static b200neox *g_model = nullptr;

bool b200_dev_open(const char *model_path, int n_ctx)
{
    const char *e = getenv("B200NEOX");
    if (e && atoi(e) == 0) return false;
    const char *dev = getenv("B200Q4_DEVICE");
    if (b200q4_init(dev ? atoi(dev) : 0) != B200Q4_OK) {
        fprintf(stderr, "b200_dev_open: %s\n", b200q4_last_error());
        exit(1);                                           // no CPU fallback behind this binding
    }
    g_model = b200neox_load(model_path, n_ctx, 64);
    if (!g_model) {
        fprintf(stderr, "b200_dev_open: %s\n", b200neox_last_error());
        exit(1);
    }
    return true;
}

// gptneox_eval's contract (vsim.cpp:470-476, 722-746): logits of the LAST token of the batch in embd_w; the first call
// reports the memory one token needs (the device graph has no per-token host arena: any non-zero value ends the probe)
bool b200_dev_eval(int n_threads, int n_past, const int *tokens, int N, std::vector<float> &embd_w, size_t &mem_per_token, int n_vocab)
{
    if (!g_model) return false;
    embd_w.resize((size_t)n_vocab);
    if (b200neox_eval(g_model, tokens, N, n_past, n_threads, embd_w.data()) != B200Q4_OK) {
        fprintf(stderr, "b200_dev_eval: %s\n", b200neox_last_error());
        exit(1);
    }
    if (mem_per_token == 0) mem_per_token = 1;
    return true;
}
