// TEST INFRASTRUCTURE ONLY: sims/balls on the reference CPU backend
// (sphere-sphere / sphere-plane / sphere-hull contacts through the reference's
// own narrowphase + GJK + XPBD).
#include <madrona/mw_cpu.hpp>
#include "../sims/balls/sim.hpp"
#include "harness.hpp"

using namespace balls;

static madrona::phys::ObjectManager *loadObjects(const char *path)
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
    return (madrona::phys::ObjectManager *)blob;
}

int main(int argc, char **argv)
{
    oracle::Args args = oracle::parseArgs(argc, argv);
    const char *objects_path = nullptr;
    for (int i = 1; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "--objects")) objects_path = argv[i + 1];
    }
    if (!objects_path) {
        fprintf(stderr, "--objects <blob> required\n");
        return 1;
    }
    Config cfg { loadObjects(objects_path) };
    std::vector<WorldInit> inits(args.numWorlds);
    for (int64_t i = 0; i < args.numWorlds; i++) inits[i].seed = (uint32_t)(args.extra[0] + i);

    using Exec = madrona::TaskGraphExecutor<Engine, Sim, Config, WorldInit>;
    Exec exec({
        .numWorlds = (uint32_t)args.numWorlds,
        .numExportedBuffers = (uint32_t)ExportID::NumExports,
        .numWorkers = (uint32_t)args.numWorkers,
    }, cfg, inits.data(), (madrona::CountT)TaskGraphID::NumTaskGraphs);

    const size_t bodies = (size_t)args.numWorlds * kMaxBodies;
    return oracle::runTrace(exec, args, {},
        { { (int)ExportID::BodyPos, [=] { return bodies * 12; } },
          { (int)ExportID::BodyRot, [=] { return bodies * 16; } },
          { (int)ExportID::BodyVel, [=] { return bodies * 24; } },
          { (int)ExportID::BodyEntity, [=] { return bodies * 8; } } });
}
