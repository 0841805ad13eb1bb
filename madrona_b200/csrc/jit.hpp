#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include <unistd.h>

namespace mb2 {

struct JitModule {
    std::vector<char> cubin;
    // nodeKernels[i] is the mangled name of a mwGPU::nodeKern<NodeT>
    // instantiation, nodeMetas[i] the matching mwGPU::nodeMeta<NodeT> global.
    std::vector<std::string> nodeKernels;
    std::vector<std::string> nodeMetas;
    std::string cachePath;
    bool fromCache = false;
};

std::string moduleDir();

bool jitCompile(const std::vector<std::string> &sources,
                const std::vector<std::string> &user_flags,
                int opt_mode, JitModule *out, std::string *err);

}
