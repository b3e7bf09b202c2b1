// File-level SPZ -> PLY conversion: the reference's exported C entry point `ConvertSpzToPly`
// (source/container/src/pipeline/post_processing/spz/src/spz_converter_api.cpp:7-30) and the `.spz` branch of its
// converter binary (spz/src/main.cpp:29-35) = loadSpz (load-spz.cc:560-575: read file, gunzip with 16|MAX_WBITS,
// deserializePackedGaussians :458-490) + unpackGaussians + saveSplatToPly (:752-838).  The host side here does the file
// I/O, the inflate (zlib, as the reference) and the header checks; the arithmetic -- every output float -- is
// spz_unpack_kernel (spz.cu) on the GPU.  This is the one entry of the library that owns device memory: it is a
// self-contained utility with no caller-side buffers to size.
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct SpzHeader {            // load-spz.cc:131-139
    uint32_t magic, version, num_points;
    uint8_t sh_degree, fractional_bits, flags, reserved;
};

bool read_file(const char* path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && fread(out->data(), 1, out->size(), f) == out->size();
    fclose(f);
    return ok;
}

bool gunzip(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
    z_stream s = {};
    s.next_in = const_cast<Bytef*>(in.data());
    s.avail_in = (uInt)in.size();
    if (inflateInit2(&s, 16 | MAX_WBITS) != Z_OK) return false;
    out->clear();
    out->resize(in.size() * 4 + 4096);
    bool ok = false;
    for (;;) {
        if (s.total_out == out->size()) out->resize(out->size() * 2);
        s.next_out = out->data() + s.total_out;
        s.avail_out = (uInt)std::min<size_t>(out->size() - s.total_out, 1u << 30);
        const int r = inflate(&s, Z_NO_FLUSH);
        if (r == Z_STREAM_END) { ok = true; break; }
        if (r != Z_OK) break;
    }
    out->resize(s.total_out);
    inflateEnd(&s);
    return ok;
}

struct DeviceBuffer {         // freed on every return path
    void* p = nullptr;
    ~DeviceBuffer() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" int b2s_convert_spz_to_ply(const char* spz_path, const char* ply_path) {
    B2S_REQUIRE(spz_path && ply_path, "null path");
    std::vector<uint8_t> gz, raw;
    if (!read_file(spz_path, &gz)) {
        b2s::set_error("%s: SPZ file does not exist or cannot be read: %s", __func__, spz_path);
        return B2S_E_BADARG;
    }
    if (!gunzip(gz, &raw) || raw.size() < sizeof(SpzHeader)) {
        b2s::set_error("%s: %s is not a gzip container with an SPZ header", __func__, spz_path);
        return B2S_E_BADARG;
    }
    SpzHeader h;
    memcpy(&h, raw.data(), sizeof(h));
    B2S_REQUIRE(h.magic == 0x5053474eu, "header not found (magic)");                      // load-spz.cc:463
    B2S_REQUIRE(h.version >= 1 && h.version <= 2, "version not supported");                // :467
    B2S_REQUIRE(h.num_points <= 10000000u, "too many points");                             // :471
    B2S_REQUIRE(h.sh_degree <= 3, "unsupported SH degree");                                // :475
    if (h.version == 1) {
        b2s::set_error("%s: float16 positions (SPZ version 1, never released) are not supported", __func__);
        return B2S_E_UNSUPPORTED;
    }
    static const int kShDim[4] = {0, 3, 8, 15};                                            // :56-70 dimForDegree
    const int64_t n = h.num_points;
    const int sh_dim = kShDim[h.sh_degree];
    const int64_t in_bytes = n * (9 + 1 + 3 + 3 + 3 + 3 * sh_dim), D = 17 + 3 * sh_dim;
    B2S_REQUIRE((int64_t)raw.size() - (int64_t)sizeof(h) >= in_bytes, "read error (payload shorter than the header says)");  // :497
    B2S_REQUIRE(n > 0, "empty cloud (the reference divides by zero here, load-spz.cc:760)");

    std::vector<float> rows((size_t)(n * D));
    {
        DeviceBuffer d_in, d_out;
        B2S_CUDA(cudaMalloc(&d_in.p, (size_t)in_bytes));
        B2S_CUDA(cudaMalloc(&d_out.p, rows.size() * sizeof(float)));
        B2S_CUDA(cudaMemcpy(d_in.p, raw.data() + sizeof(h), (size_t)in_bytes, cudaMemcpyHostToDevice));
        const uint8_t* b = static_cast<const uint8_t*>(d_in.p);                            // serialised order, :448-453
        const uint8_t *pos = b, *alp = pos + 9 * n, *col = alp + n, *sca = col + 3 * n, *rot = sca + 3 * n, *sh = rot + 3 * n;
        const int rc = b2s_spz_unpack(n, sh_dim, h.fractional_bits, pos, alp, col, sca, rot, sh_dim ? sh : nullptr,
                                      static_cast<float*>(d_out.p), nullptr);
        if (rc != 0) return rc;
        B2S_CUDA(cudaMemcpy(rows.data(), d_out.p, rows.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }

    std::string head = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(n) + "\n";   // :795-820
    for (const char* k : {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"}) head += std::string("property float ") + k + "\n";
    for (int i = 0; i < 3 * sh_dim; ++i) head += "property float f_rest_" + std::to_string(i) + "\n";
    for (const char* k : {"opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"}) head += std::string("property float ") + k + "\n";
    head += "end_header\n";
    FILE* f = fopen(ply_path, "wb");
    if (!f) {
        b2s::set_error("%s: unable to open for writing: %s", __func__, ply_path);
        return B2S_E_BADARG;
    }
    const bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() &&
                    fwrite(rows.data(), sizeof(float), rows.size(), f) == rows.size();
    if (fclose(f) != 0 || !ok) {
        b2s::set_error("%s: failed to write to: %s", __func__, ply_path);
        return B2S_E_BADARG;
    }
    return 0;
}

// The reference's own exported name and signature (spz_converter_api.cpp:7): true on success, message on stderr otherwise.
extern "C" B2S_API bool ConvertSpzToPly(const char* spzFilePath, const char* plyFilePath) {
    if (b2s_convert_spz_to_ply(spzFilePath, plyFilePath) == 0) return true;
    fprintf(stderr, "Error: %s\n", b2s_last_error());
    return false;
}
