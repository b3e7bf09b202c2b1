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

// ----------------------------------------------------------------------------------------------------------
// PLY -> SPZ: the `.ply` branch of the reference's converter (spz/src/main.cpp:22-28) = loadSplatFromPly
// (load-spz.cc:568-750) + saveSpz (:492-500, :546-558: packGaussians, serialise, gzip).  The host parses the header exactly
// as the reference does (the same two fixed lines, `comment` lines skipped, float properties only, fields found by name),
// uploads the vertex table as it lies in the file, and the GPU does the rest: a gather kernel splits the interleaved rows
// into the codec's arrays (the [N,C,S] -> [N,S,C] turn of the SH coefficients included, :742-747) and spz_pack_kernel
// quantises them.  The gzip container is written with the reference's parameters (:189-191), so the file is byte-identical.
// ----------------------------------------------------------------------------------------------------------
namespace b2s {
// one thread per (row, slot): slots = x y z | scale_0..2 | rot_0..3 (w first) | opacity | f_dc_0..2 | sh[j][c]
__global__ void __launch_bounds__(256) ply_gather_kernel(int64_t N, int F, int sh_dim, const float* __restrict__ table,
                                                         const int* __restrict__ cols, float* __restrict__ means,
                                                         float* __restrict__ scales, float* __restrict__ quats,
                                                         float* __restrict__ opac, float* __restrict__ sh0, float* __restrict__ shN) {
    const int S = 14 + 3 * sh_dim;
    const int64_t total = N * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / S;
        const int s = (int)(e - i * S);
        const float v = table[i * F + cols[s]];
        if (s < 3) means[3 * i + s] = v;
        else if (s < 6) scales[3 * i + (s - 3)] = v;
        else if (s < 10) quats[4 * i + (s - 6)] = v;
        else if (s == 10) opac[i] = v;
        else if (s < 14) sh0[3 * i + (s - 11)] = v;
        else shN[i * (3 * sh_dim) + (s - 14)] = v;
    }
}
}  // namespace b2s

namespace {

bool next_line(const std::vector<uint8_t>& buf, size_t* pos, std::string* line) {      // std::getline on the header
    if (*pos >= buf.size()) return false;
    size_t e = *pos;
    while (e < buf.size() && buf[e] != '\n') ++e;
    line->assign(reinterpret_cast<const char*>(buf.data()) + *pos, e - *pos);
    *pos = e < buf.size() ? e + 1 : e;
    return true;
}

bool gzip_to(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {               // load-spz.cc:182-211
    z_stream s = {};
    if (deflateInit2(&s, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 16 + MAX_WBITS, 9, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out->resize(deflateBound(&s, (uLong)in.size()) + 64);
    s.next_in = const_cast<Bytef*>(in.data());
    s.avail_in = (uInt)in.size();
    s.next_out = out->data();
    s.avail_out = (uInt)out->size();
    const int r = deflate(&s, Z_FINISH);
    out->resize(s.total_out);
    deflateEnd(&s);
    return r == Z_STREAM_END;
}

}  // namespace

extern "C" int b2s_convert_ply_to_spz(const char* ply_path, const char* spz_path) {
    B2S_REQUIRE(ply_path && spz_path, "null path");
    std::vector<uint8_t> file;
    if (!read_file(ply_path, &file)) {
        b2s::set_error("%s: unable to open: %s", __func__, ply_path);
        return B2S_E_BADARG;
    }
    size_t pos = 0;
    std::string line;
    B2S_REQUIRE(next_line(file, &pos, &line) && line == "ply", "not a .ply file");                                   // :582
    B2S_REQUIRE(next_line(file, &pos, &line) && line == "format binary_little_endian 1.0", "unsupported .ply format");  // :589
    long long n = -1;
    while (next_line(file, &pos, &line)) {                                                                           // :597-615
        if (line.compare(0, 7, "comment") == 0) continue;
        if (line.compare(0, 15, "element vertex ") == 0) {
            char* end = nullptr;
            n = strtoll(line.c_str() + 14, &end, 10);
            if (end == line.c_str() + 14) n = -1;
            break;
        }
    }
    B2S_REQUIRE(n > 0 && n <= 10ll * 1024 * 1024, "invalid vertex count");                                            // :617
    std::vector<std::string> names;
    for (;;) {                                                                                                       // :627-645
        B2S_REQUIRE(next_line(file, &pos, &line), "header without end_header");
        if (line.compare(0, 7, "comment") == 0) continue;
        if (line == "end_header") break;
        B2S_REQUIRE(line.compare(0, 15, "property float ") == 0, "unsupported property data type");
        names.push_back(line.substr(15));
    }
    const int F = (int)names.size();
    auto index = [&](const std::string& name) {
        for (int i = F - 1; i >= 0; --i)          // a repeated name keeps its last index, as the reference's map does
            if (names[i] == name) return i;
        return -1;
    };
    std::vector<int> cols;
    for (const char* k : {"x", "y", "z", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "opacity", "f_dc_0",
                          "f_dc_1", "f_dc_2"}) {
        const int c = index(k);
        if (c < 0) {
            b2s::set_error("%s: missing field: %s", __func__, k);                                                    // :651
            return B2S_E_BADARG;
        }
        cols.push_back(c);
    }
    std::vector<int> sh_idx;                                                                                         // :692-699
    for (int i = 0; i < 45; ++i) {
        const int c = index("f_rest_" + std::to_string(i));
        if (c < 0) break;
        sh_idx.push_back(c);
    }
    const int sh_dim = (int)sh_idx.size() / 3;
    // the reference goes on with degreeForDim(shDim) and then fails its size check in saveSpz unless shDim is 0, 3, 8 or 15
    B2S_REQUIRE((sh_dim == 0 || sh_dim == 3 || sh_dim == 8 || sh_dim == 15) && (int)sh_idx.size() == 3 * sh_dim,
                "f_rest_* must hold 0, 9, 24 or 45 coefficients");
    for (int j = 0; j < sh_dim; ++j)
        for (int c = 0; c < 3; ++c) cols.push_back(sh_idx[c * sh_dim + j]);                                          // [N,C,S] -> [N,S,C]
    const size_t table_bytes = (size_t)n * F * sizeof(float);
    B2S_REQUIRE(file.size() - pos >= table_bytes, "unable to load data (file shorter than its header says)");        // :703
    static const int kDegree[16] = {0, 0, 0, 1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 3};

    const int S = 14 + 3 * sh_dim;
    const size_t packed_bytes = (size_t)n * (9 + 1 + 3 + 3 + 3 + 3 * sh_dim);
    std::vector<uint8_t> payload(sizeof(SpzHeader) + packed_bytes);
    {
        DeviceBuffer d_table, d_cols, d_arrays, d_packed;
        B2S_CUDA(cudaMalloc(&d_table.p, table_bytes));
        B2S_CUDA(cudaMalloc(&d_cols.p, cols.size() * sizeof(int)));
        B2S_CUDA(cudaMalloc(&d_arrays.p, (size_t)n * S * sizeof(float)));
        B2S_CUDA(cudaMalloc(&d_packed.p, packed_bytes));
        B2S_CUDA(cudaMemcpy(d_table.p, file.data() + pos, table_bytes, cudaMemcpyHostToDevice));
        B2S_CUDA(cudaMemcpy(d_cols.p, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice));
        float* a = static_cast<float*>(d_arrays.p);
        float *means = a, *scales = means + 3 * n, *quats = scales + 3 * n, *opac = quats + 4 * n, *sh0 = opac + n, *shN = sh0 + 3 * n;
        const int64_t total = n * S;
        b2s::ply_gather_kernel<<<(unsigned)std::min<int64_t>(b2s::cdiv(total, 256), 148 * 64), 256>>>(
            n, F, sh_dim, static_cast<const float*>(d_table.p), static_cast<const int*>(d_cols.p), means, scales, quats, opac, sh0, shN);
        B2S_LAUNCH_CHECK();
        uint8_t* b = static_cast<uint8_t*>(d_packed.p);                                    // serialised order, :448-453
        uint8_t *o_pos = b, *o_alp = o_pos + 9 * n, *o_col = o_alp + n, *o_sca = o_col + 3 * n, *o_rot = o_sca + 3 * n, *o_sh = o_rot + 3 * n;
        const int rc = b2s_spz_pack(n, sh_dim, means, scales, quats, opac, sh0, sh_dim ? shN : nullptr, sh_dim, o_pos, o_alp, o_col,
                                    o_sca, o_rot, sh_dim ? o_sh : nullptr, nullptr);
        if (rc != 0) return rc;
        B2S_CUDA(cudaMemcpy(payload.data() + sizeof(SpzHeader), d_packed.p, packed_bytes, cudaMemcpyDeviceToHost));
    }
    SpzHeader h = {0x5053474eu, 2u, (uint32_t)n, (uint8_t)kDegree[sh_dim], 12, 0, 0};                                 // :131-139, :443-447
    memcpy(payload.data(), &h, sizeof(h));
    std::vector<uint8_t> gz;
    if (!gzip_to(payload, &gz)) {
        b2s::set_error("%s: gzip failed", __func__);
        return B2S_E_BADARG;
    }
    FILE* f = fopen(spz_path, "wb");
    if (!f) {
        b2s::set_error("%s: unable to open for writing: %s", __func__, spz_path);
        return B2S_E_BADARG;
    }
    const bool ok = fwrite(gz.data(), 1, gz.size(), f) == gz.size();
    if (fclose(f) != 0 || !ok) {
        b2s::set_error("%s: failed to write to: %s", __func__, spz_path);
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
