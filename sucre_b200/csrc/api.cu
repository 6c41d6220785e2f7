// ABI bookkeeping (version, thread-local error string) and the footprint upload of a host scene.
#include <string>

#include "common.cuh"

namespace sucre {
static thread_local std::string g_error;

int set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}
void clear_error() { g_error.clear(); }
}  // namespace sucre

extern "C" int sucre_abi_version(void) { return SUCRE_ABI_VERSION; }
extern "C" const char* sucre_last_error(void) { return sucre::g_error.c_str(); }

extern "C" int sucre_scene_upload(void* dst, const void* src_host, int n, const int32_t* src_index_host, int width, int height,
                                  int pixel_bytes, const int32_t* rects_host, int64_t* copied_bytes_host, void* stream) {
    using namespace sucre;
    clear_error();
    SUCRE_REQUIRE(dst && src_host && src_index_host && rects_host, "sucre_scene_upload: null pointer");
    SUCRE_REQUIRE(n >= 0 && width > 0 && height > 0 && pixel_bytes > 0, "sucre_scene_upload: bad sizes");
    const size_t row_bytes = (size_t)width * pixel_bytes, view_bytes = row_bytes * height;
    int64_t copied = 0;
    for (int i = 0; i < n; ++i) {
        const int32_t* r = rects_host + 4 * i;
        const int x0 = r[0], y0 = r[1], x1 = r[2], y1 = r[3];
        if (x1 <= x0 || y1 <= y0) continue;
        SUCRE_REQUIRE(x0 >= 0 && y0 >= 0 && x1 <= width && y1 <= height && src_index_host[i] >= 0,
                      "sucre_scene_upload: rectangle %d = [%d, %d) x [%d, %d) outside the %d x %d image", i, x0, x1, y0, y1, width, height);
        const size_t off = (size_t)y0 * row_bytes + (size_t)x0 * pixel_bytes;
        char* d = static_cast<char*>(dst) + (size_t)i * view_bytes + off;
        const char* s = static_cast<const char*>(src_host) + (size_t)src_index_host[i] * view_bytes + off;
        const size_t w = (size_t)(x1 - x0) * pixel_bytes, h = (size_t)(y1 - y0);
        if (w == row_bytes) {  // whole rows: one contiguous block
            SUCRE_CUDA(cudaMemcpyAsync(d, s, w * h, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        } else {
            SUCRE_CUDA(cudaMemcpy2DAsync(d, row_bytes, s, row_bytes, w, h, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        }
        copied += (int64_t)(w * h);
    }
    if (copied_bytes_host) *copied_bytes_host = copied;
    return 0;
}
