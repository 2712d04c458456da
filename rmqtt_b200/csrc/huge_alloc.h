// Allocator for the host mirrors of the big device tables (multi-GB open-addressing tables that the mutation path
// probes at random): large blocks come from anonymous mmap with MADV_HUGEPAGE, so a random probe costs one
// translation out of a 2-MiB-page TLB instead of a 4-level walk per 4-KiB page.  Purely a host-side speed-up of
// gm_sub_add / gm_bulk_load / gm_compact; a kernel without transparent huge pages simply ignores the advice.
#pragma once
#include <sys/mman.h>

#include <cstddef>
#include <cstdlib>
#include <new>

namespace gm {

template <class T>
struct HugeAlloc {
    using value_type = T;
    static constexpr size_t kHugeThreshold = size_t(4) << 20;
    static constexpr size_t kAlign = alignof(T) > 64 ? alignof(T) : 64;
    HugeAlloc() = default;
    template <class U> HugeAlloc(const HugeAlloc<U>&) {}
    static size_t rounded(size_t bytes) { return (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1); }
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes >= kHugeThreshold) {
            void* p = mmap(nullptr, rounded(bytes), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) throw std::bad_alloc();
            madvise(p, rounded(bytes), MADV_HUGEPAGE);
            return static_cast<T*>(p);
        }
        void* p = std::aligned_alloc(kAlign, (bytes + kAlign - 1) / kAlign * kAlign + kAlign);
        if (!p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t n) noexcept {
        const size_t bytes = n * sizeof(T);
        if (bytes >= kHugeThreshold) munmap(p, rounded(bytes));
        else std::free(p);
    }
    template <class U> bool operator==(const HugeAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const HugeAlloc<U>&) const { return false; }
};

}  // namespace gm
