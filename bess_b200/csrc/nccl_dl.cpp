#include "nccl_dl.h"

#include <dlfcn.h>

#include <mutex>

#include "engine.h"

namespace bess {

const NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            err = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        auto sym = [&](const char *name) {
            void *p = dlsym(h, name);
            if (!p && err.empty()) err = std::string("libnccl: missing symbol ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!err.empty()) throw EngineError{err};
    return api;
}

}  // namespace bess
