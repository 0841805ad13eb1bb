// TEST INFRASTRUCTURE ONLY (see oracle/Makefile header).
// Drives a fixture simulator on the *reference* CPU backend
// (madrona::TaskGraphExecutor, include/madrona/mw_cpu.hpp:73-110) and dumps the
// exported columns after every step, so tests can compare the B200 engine
// against the reference's own execution of the same simulator sources.
//
// Trace protocol (little endian, raw):
//   input  file: per step, for each input slot in order:  numWorlds * bytesPerWorld bytes
//   output file: after init ("step -1") and after every step, for each output
//                slot in order: u64 numBytes, then the bytes.
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <functional>

namespace oracle {

struct InSlot {
    int slot;
    size_t bytesPerWorld;
};

struct OutSlot {
    int slot;
    // total bytes to dump for this slot right now (fixed tables: W * bytes)
    std::function<size_t()> numBytes;
};

struct Args {
    int64_t numWorlds = 4;
    int64_t numSteps = 10;
    int64_t numWorkers = 1;
    const char *inPath = nullptr;
    const char *outPath = nullptr;
    int64_t extra[4] = { 0, 0, 0, 0 };
};

inline Args parseArgs(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc; i++) {
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : "0"; };
        if (!strcmp(argv[i], "--worlds")) a.numWorlds = atoll(next());
        else if (!strcmp(argv[i], "--steps")) a.numSteps = atoll(next());
        else if (!strcmp(argv[i], "--workers")) a.numWorkers = atoll(next());
        else if (!strcmp(argv[i], "--in")) a.inPath = next();
        else if (!strcmp(argv[i], "--out")) a.outPath = next();
        else if (!strcmp(argv[i], "--x0")) a.extra[0] = atoll(next());
        else if (!strcmp(argv[i], "--x1")) a.extra[1] = atoll(next());
        else if (!strcmp(argv[i], "--x2")) a.extra[2] = atoll(next());
        else if (!strcmp(argv[i], "--x3")) a.extra[3] = atoll(next());
    }
    return a;
}

// ObjectManager blob written by sims/objects.py:write_blob_file (u64 size,
// u64 numRelocs, relocs[], blob): loaded and relocated to host addresses.
inline void *loadObjectsBlob(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "cannot open objects blob %s\n", path);
        exit(1);
    }
    uint64_t size = 0, num_relocs = 0;
    if (fread(&size, 8, 1, f) != 1 || fread(&num_relocs, 8, 1, f) != 1) exit(1);
    std::vector<uint64_t> relocs(num_relocs);
    if (num_relocs && fread(relocs.data(), 8, num_relocs, f) != num_relocs) exit(1);
    char *blob = (char *)aligned_alloc(64, (size + 63) / 64 * 64);
    if (fread(blob, 1, size, f) != size) exit(1);
    fclose(f);
    for (uint64_t where : relocs) {
        uint64_t off;
        memcpy(&off, blob + where, 8);
        uint64_t addr = (uint64_t)(uintptr_t)blob + off;
        memcpy(blob + where, &addr, 8);
    }
    return blob;
}

inline const char *objectsPathArg(int argc, char **argv)
{
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "--objects")) return argv[i + 1];
    }
    fprintf(stderr, "--objects <blob> required\n");
    exit(1);
}

template <typename ExecT>
int runTrace(ExecT &exec, const Args &args,
             const std::vector<InSlot> &ins, const std::vector<OutSlot> &outs)
{
    FILE *fin = args.inPath ? fopen(args.inPath, "rb") : nullptr;
    FILE *fout = args.outPath ? fopen(args.outPath, "wb") : nullptr;
    if (args.inPath && !fin) {
        fprintf(stderr, "cannot open %s\n", args.inPath);
        return 1;
    }

    auto dump = [&]() {
        if (!fout) return;
        for (const OutSlot &o : outs) {
            uint64_t n = o.numBytes();
            fwrite(&n, sizeof(n), 1, fout);
            if (n) fwrite(exec.getExported(o.slot), 1, n, fout);
        }
    };
    dump();

    double run_seconds = 0;
    for (int64_t step = 0; step < args.numSteps; step++) {
        if (fin) {
            for (const InSlot &in : ins) {
                size_t n = in.bytesPerWorld * (size_t)args.numWorlds;
                if (fread(exec.getExported(in.slot), 1, n, fin) != n) {
                    fprintf(stderr, "short read in input trace at step %ld\n", (long)step);
                    return 1;
                }
            }
        }
        auto t0 = std::chrono::steady_clock::now();
        exec.run();
        auto t1 = std::chrono::steady_clock::now();
        run_seconds += std::chrono::duration<double>(t1 - t0).count();
        dump();
    }
    if (fin) fclose(fin);
    if (fout) fclose(fout);
    // machine-readable timing line for bench.py's cpu_baseline
    printf("{\"worlds\": %ld, \"steps\": %ld, \"workers\": %ld, \"seconds\": %.6f, "
           "\"steps_per_sec\": %.3f}\n",
           (long)args.numWorlds, (long)args.numSteps, (long)args.numWorkers, run_seconds,
           run_seconds > 0 ? (double)args.numWorlds * args.numSteps / run_seconds : 0.0);
    return 0;
}

}
