#include "multi_device.h"

#include <dlfcn.h>
#include <nccl.h>

namespace kgpu {

struct MultiDevice::Impl {
    void *dl = nullptr;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

MultiDevice *MultiDevice::create(const std::vector<int> &devs, std::string *why) {
    Impl *im = new Impl();
    im->devs = devs;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        im->dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (im->dl) break;
    }
    if (!im->dl) {
        *why = std::string("multi-device handle needs NCCL: dlopen(libnccl.so.2) failed: ") + dlerror();
        delete im;
        return nullptr;
    }
#define KGPU_SYM(field, name)                                              \
    im->field = reinterpret_cast<decltype(im->field)>(dlsym(im->dl, name)); \
    if (!im->field) { *why = std::string("NCCL symbol missing: ") + name; delete im; return nullptr; }
    KGPU_SYM(CommInitAll, "ncclCommInitAll")
    KGPU_SYM(CommDestroy, "ncclCommDestroy")
    KGPU_SYM(AllGather, "ncclAllGather")
    KGPU_SYM(GroupStart, "ncclGroupStart")
    KGPU_SYM(GroupEnd, "ncclGroupEnd")
    KGPU_SYM(GetErrorString, "ncclGetErrorString")
#undef KGPU_SYM
    im->comms.resize(devs.size());
    ncclResult_t r = im->CommInitAll(im->comms.data(), (int)devs.size(), devs.data());
    if (r != ncclSuccess) {
        *why = std::string("ncclCommInitAll failed: ") + im->GetErrorString(r);
        im->comms.clear();
        delete im;
        return nullptr;
    }
    MultiDevice *m = new MultiDevice();
    m->impl_ = im;
    return m;
}

MultiDevice::~MultiDevice() {
    if (!impl_) return;
    for (size_t i = 0; i < impl_->comms.size(); i++) {
        cudaSetDevice(impl_->devs[i]);
        impl_->CommDestroy(impl_->comms[i]);
    }
    delete impl_;
}

bool MultiDevice::all_gather_u64(const std::vector<const void *> &send, const std::vector<void *> &recv, size_t count,
                                 const std::vector<cudaStream_t> &streams, std::string *why) {
    ncclResult_t r = impl_->GroupStart();
    for (size_t i = 0; r == ncclSuccess && i < impl_->comms.size(); i++)
        r = impl_->AllGather(send[i], recv[i], count, ncclUint64, impl_->comms[i], streams[i]);
    ncclResult_t e = impl_->GroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) {
        *why = std::string("ncclAllGather failed: ") + impl_->GetErrorString(r);
        return false;
    }
    return true;
}

}  // namespace kgpu
