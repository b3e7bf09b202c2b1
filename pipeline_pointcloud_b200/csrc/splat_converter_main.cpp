// `splat_converter <file>` -- the command line of the reference's converter binary (spz/src/main.cpp), which the Spz-Export
// component runs as `splat_converter DIR/splat.ply` (source/container/src/main.py:1532-1545): a `.ply` becomes a `.spz` next to
// it, a `.spz` a `.ply`; anything else is an error.  Same messages, same exit codes (0 / 1); the work is libb2s.so's
// (b2s_convert_ply_to_spz / b2s_convert_spz_to_ply: host parse + gzip, arithmetic on the GPU).
#include <sys/stat.h>

#include <cstdio>
#include <string>

#include "../../include/b2s.h"

int main(int argc, char* argv[]) {
    if (argc != 2) {
        fprintf(stderr, "Usage: %s <input_file>\n", argv[0]);
        return 1;
    }
    const std::string in = argv[1];
    struct stat st;
    if (stat(in.c_str(), &st) != 0) {
        fprintf(stderr, "Error: File does not exist: %s\n", in.c_str());
        return 1;
    }
    const size_t slash = in.find_last_of('/');
    const size_t dot = in.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash + 1);   // ".ply" alone is a stem
    const std::string ext = has_ext ? in.substr(dot) : "";
    const std::string stem = has_ext ? in.substr(0, dot) : in;
    std::string out;
    int rc;
    if (ext == ".ply") {
        out = stem + ".spz";
        rc = b2s_convert_ply_to_spz(in.c_str(), out.c_str());
        if (rc != 0) fprintf(stderr, "Error: Failed to save SPZ file (%s)\n", b2s_last_error());
    } else if (ext == ".spz") {
        out = stem + ".ply";
        rc = b2s_convert_spz_to_ply(in.c_str(), out.c_str());
        if (rc != 0) fprintf(stderr, "Error: Failed to save PLY file (%s)\n", b2s_last_error());
    } else {
        fprintf(stderr, "Error: Unsupported file extension: %s\n", ext.c_str());
        return 1;
    }
    if (rc != 0) return 1;
    printf("Converted %s to \"%s\"\n", in.c_str(), out.c_str());     // the reference streams a std::filesystem::path: quoted
    return 0;
}
