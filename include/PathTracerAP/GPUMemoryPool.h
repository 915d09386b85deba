// GPUMemoryPool.h - facade of the reference's GPUMemoryPool<T> (GPUMemoryPool.h:10-46).
// The reference allocates one cudaMallocManaged block per array plus a managed copy of the pool object; libptap keeps all device
// memory in its own arena, so this class is only the HOST-VISIBLE view the reference exposes through RenderData: `size` elements
// at `pool`.  Renderer fills the image pool after renderLoop(); allocate()/free() manage plain host memory.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>

template <typename T>
class GPUMemoryPool {
public:
    GPUMemoryPool() : size(0), pool(nullptr) {}
    GPUMemoryPool* getInstance() { return this; }
    void allocate(const std::vector<T>& data)
    {
        free();
        size = (int)data.size();
        pool = static_cast<T*>(std::malloc(sizeof(T) * (data.empty() ? 1 : data.size())));
        if (!data.empty()) std::memcpy(pool, data.data(), sizeof(T) * data.size());
    }
    void free() { std::free(pool); pool = nullptr; size = 0; }

    int size;
    T* pool;
};
