// alignedallocator.hpp -- std-compatible aligned allocator (interface of reference
// include/alignedallocator.hpp).  Only used for the (unused) host `state_storage` member.
#ifndef IQS_ALIGNED_ALLOCATOR_HPP
#define IQS_ALIGNED_ALLOCATOR_HPP
#include <cstddef>
#include <cstdlib>
#include <new>
namespace iqs {
template <typename T, unsigned int Alignment>
class AlignedAllocator {
 public:
  typedef T value_type;
  typedef T *pointer;
  typedef T const *const_pointer;
  typedef T &reference;
  typedef T const &const_reference;
  typedef std::size_t size_type;
  typedef std::ptrdiff_t difference_type;
  template <typename U>
  struct rebind {
    typedef AlignedAllocator<U, Alignment> other;
  };
  AlignedAllocator() noexcept {}
  AlignedAllocator(AlignedAllocator const &) noexcept {}
  template <typename U>
  AlignedAllocator(AlignedAllocator<U, Alignment> const &) noexcept {}
  pointer allocate(size_type n) {
    void *p = nullptr;
    if (n == 0) return nullptr;
    if (posix_memalign(&p, Alignment < sizeof(void *) ? sizeof(void *) : Alignment, n * sizeof(T)) != 0) throw std::bad_alloc();
    return static_cast<pointer>(p);
  }
  void deallocate(pointer p, size_type) noexcept { std::free(p); }
  size_type max_size() const noexcept { return static_cast<size_type>(-1) / sizeof(T); }
  template <class U, class... Args>
  void construct(U *p, Args &&...args) { ::new ((void *)p) U(static_cast<Args &&>(args)...); }
  template <class U>
  void destroy(U *p) { p->~U(); }
  bool operator==(AlignedAllocator const &) const { return true; }
  bool operator!=(AlignedAllocator const &) const { return false; }
};
}  // namespace iqs
#endif
