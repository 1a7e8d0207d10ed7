// r3d_textin.cu -- a8: the point-cloud text readers of the OctoMap scripts on a host thread pool (host code only).
//
// txt_read of octomap/txt_transfer_octomap.py:16-28 parses every `x,y,z` line with float() in a Python loop, the PLY
// twin (octomap/ply_transfer_octomap.py:16-40) skips 8 lines, splits on whitespace and stops after point 5 400 000.
// Here the file is split at line boundaries into one piece per worker; every piece is parsed (r3d_strtod.cuh: correctly
// rounded like Python's float(), strtod for whatever is not a plain decimal literal) once, and the pieces' rows are then
// copied to their final offsets, so the points come out in file order.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "r3d_common.cuh"
#include "r3d_strtod.cuh"

namespace {

using r3d::set_error;

inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }

// [sign] (digits [. [digits]] | . digits) [(e|E) [sign] digits]: the decimal part of Python's float() grammar
bool decimal_syntax(const char* p, const char* e) {
    if (p < e && (*p == '+' || *p == '-')) ++p;
    int nd = 0;
    while (p < e && is_digit(*p)) { ++p; ++nd; }
    if (p < e && *p == '.') {
        ++p;
        while (p < e && is_digit(*p)) { ++p; ++nd; }
    }
    if (nd == 0) return false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        if (p < e && (*p == '+' || *p == '-')) ++p;
        if (p >= e || !is_digit(*p)) return false;
        while (p < e && is_digit(*p)) ++p;
    }
    return p == e;
}

// One field (no surrounding blanks) with float()'s verdict: the value, or false where float() raises ValueError.
bool parse_field(const char* p, const char* e, double* v) {
    if (r3d::parse_double(p, e, v)) return true;            // plain decimal literal of <= 19 digits, normal result
    if (p >= e) return false;
    // inf / infinity / nan, any case, optional sign
    {
        const char* q = p;
        bool neg = false;
        if (*q == '+' || *q == '-') { neg = *q == '-'; ++q; }
        const size_t n = (size_t)(e - q);
        auto ieq = [&](const char* w) { if (strlen(w) != n) return false; for (size_t i = 0; i < n; ++i) if ((q[i] | 0x20) != w[i]) return false; return true; };
        if (ieq("inf") || ieq("infinity")) { *v = neg ? -HUGE_VAL : HUGE_VAL; return true; }
        if (ieq("nan")) { *v = neg ? -NAN : NAN; return true; }
    }
    // PEP 515 underscores (only between digits) are dropped; what remains must be a decimal literal -- strtod would also
    // take hexadecimal floats and "nan(...)", which float() rejects
    std::string tmp;
    tmp.reserve((size_t)(e - p));
    for (const char* q = p; q < e; ++q) {
        if (*q == '_') {
            if (q == p || q + 1 == e || !is_digit(q[-1]) || !is_digit(q[1])) return false;
            continue;
        }
        tmp.push_back(*q);
    }
    const char* b = tmp.c_str();
    if (!decimal_syntax(b, b + tmp.size())) return false;
    if (r3d::parse_double(b, b + tmp.size(), v)) return true;
    char* end = nullptr;
    *v = strtod(b, &end);                                    // long digit strings, subnormal / overflowing results
    return end == b + tmp.size();
}

// One line (no newline) as the reference's readers see it: `str_tofloat(line.split(','))` (comma_mode, surrounding blanks
// allowed like float(" 1.5 ")) or `str_tofloat(line.split())`: EVERY field goes through float(), the first three are the
// point.  Returns 3 for a point, 0 for an empty / blank line (skipped: the one documented deviation, the reference's own
// PLY writer ends its files with such a line), -1 where the reference raises: a field float() rejects, or fewer than three
// fields (the binding then indexes past the array).
inline int parse_row(const char* p, const char* e, bool comma_mode, double v[3]) {
    {
        const char* q = p;
        while (q < e && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        if (q == e) return 0;
    }
    int got = 0;
    while (p <= e) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
        if (!comma_mode && p == e) break;                        // split(): trailing blanks make no field
        const char* q = p;
        if (comma_mode) { while (q < e && *q != ',') ++q; }
        else { while (q < e && *q != ' ' && *q != '\t' && *q != '\r') ++q; }
        const char* te = q;
        while (te > p && (te[-1] == ' ' || te[-1] == '\t' || te[-1] == '\r')) --te;
        double tmp;
        if (te == p || !parse_field(p, te, got < 3 ? &v[got] : &tmp)) return -1;   // empty or not a number: float() raises
        ++got;
        if (comma_mode) {
            if (q >= e) break;
            p = q + 1;                                           // a trailing ',' leaves an empty last field: float('') raises
        } else {
            p = q;
        }
    }
    return got >= 3 ? 3 : -1;
}

struct Piece {
    size_t begin = 0, end = 0;   // byte range, whole lines
    uint64_t rows = 0, first = 0;
    size_t bad_at = (size_t)-1;  // byte offset of the first line the reference would raise on
    uint64_t bad_row = 0;        // points of this piece before that line
};

}  // namespace

// Points of an `x,y,z` text (comma_mode != 0) or of an ASCII PLY body (comma_mode == 0, whitespace separated, only the
// first three columns used, every column must be a number) after skipping skip_lines lines; empty / blank lines (the
// reference writer's trailing indentation) are skipped, any other line the reference would raise on is R3D_ERR_ARG with
// its line number; at most max_points points (0 = no limit: lines after the cap are not looked at, as in the reference).
// out == NULL or capacity too small: only *n_points is set (size query).  Needs no GPU.
extern "C" int r3d_read_xyz_text(const char* path, int skip_lines, int comma_mode, uint64_t max_points, double* out, uint64_t capacity,
                                 uint64_t* n_points, int n_threads) {
    if (!path || !n_points) return set_error(nullptr, R3D_ERR_ARG, "r3d_read_xyz_text: null argument");
    *n_points = 0;
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(nullptr, R3D_ERR_IO, "cannot open %s", path);
    std::vector<char> data;
    {
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n < 0) { fclose(f); return set_error(nullptr, R3D_ERR_IO, "cannot size %s", path); }
        data.resize((size_t)n);
        const size_t got = n ? fread(data.data(), 1, (size_t)n, f) : 0;
        fclose(f);
        if (got != (size_t)n) return set_error(nullptr, R3D_ERR_IO, "short read on %s", path);
    }
    const char* base = data.data();
    const size_t size = data.size();
    size_t pos = 0;
    for (int i = 0; i < skip_lines && pos < size; ++i) {
        const void* nl = memchr(base + pos, '\n', size - pos);
        pos = nl ? (size_t)((const char*)nl - base) + 1 : size;
    }
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    int nt = n_threads > 0 ? n_threads : (int)hw;
    const size_t body = size - pos;
    if ((size_t)nt > body / 65536 + 1) nt = (int)(body / 65536 + 1);
    std::vector<Piece> pieces((size_t)nt);
    size_t b = pos;
    for (int t = 0; t < nt; ++t) {
        size_t e = t + 1 == nt ? size : pos + body * (size_t)(t + 1) / (size_t)nt;
        if (e < b) e = b;
        if (e < size) {
            const void* nl = memchr(base + e, '\n', size - e);
            e = nl ? (size_t)((const char*)nl - base) + 1 : size;
        }
        pieces[(size_t)t].begin = b;
        pieces[(size_t)t].end = e;
        b = e;
    }
    const bool comma = comma_mode != 0;
    // pass 1: every piece parses its lines once; the rows are kept per piece when there is somewhere to put them
    const bool keep = out != nullptr && capacity > 0;
    std::vector<std::vector<double>> rows_of((size_t)nt);
    auto scan = [&](int t) {
        Piece& pc = pieces[(size_t)t];
        std::vector<double>& mine = rows_of[(size_t)t];
        if (keep) mine.reserve((pc.end - pc.begin) / 16 + 16);
        uint64_t rows = 0;
        size_t p = pc.begin;
        while (p < pc.end) {
            const void* nl = memchr(base + p, '\n', pc.end - p);
            const size_t le = nl ? (size_t)((const char*)nl - base) : pc.end;
            double v[3];
            const int kind = parse_row(base + p, base + le, comma, v);
            if (kind == 3) {
                if (keep) mine.insert(mine.end(), v, v + 3);
                ++rows;
            } else if (kind < 0) {
                pc.bad_at = p;
                pc.bad_row = rows;
                break;
            }
            p = le + 1;
        }
        pc.rows = rows;
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(scan, t);
        scan(0);
        for (auto& th : pool) th.join();
    }
    uint64_t total = 0;
    for (auto& pc : pieces) {
        pc.first = total;
        total += pc.rows;
        // a malformed line is an error exactly when the reference's loop would have reached it (it stops at max_points)
        if (pc.bad_at != (size_t)-1 && (!max_points || pc.first + pc.bad_row < max_points)) {
            uint64_t line = 1;
            for (size_t i = 0; i < pc.bad_at; ++i) line += base[i] == '\n';
            const void* nl = memchr(base + pc.bad_at, '\n', size - pc.bad_at);
            size_t len = nl ? (size_t)((const char*)nl - (base + pc.bad_at)) : size - pc.bad_at;
            if (len > 80) len = 80;
            return set_error(nullptr, R3D_ERR_ARG, "%s, line %llu: could not convert string to float (or fewer than three fields): '%.*s'", path,
                             (unsigned long long)line, (int)len, base + pc.bad_at);
        }
    }
    if (max_points && total > max_points) total = max_points;
    *n_points = total;
    if (!out || capacity < total) return R3D_OK;
    // pass 2: the pieces' rows to their final offsets, file order kept (pieces beyond the cap write nothing)
    {
        std::vector<std::thread> pool;
        auto fill = [&](int t) {
            const Piece& pc = pieces[(size_t)t];
            if (pc.first >= total) return;
            const uint64_t n = pc.rows < total - pc.first ? pc.rows : total - pc.first;
            if (n) memcpy(out + 3 * pc.first, rows_of[(size_t)t].data(), (size_t)n * 24);
        };
        for (int t = 1; t < nt; ++t) pool.emplace_back(fill, t);
        fill(0);
        for (auto& th : pool) th.join();
    }
    return R3D_OK;
}
