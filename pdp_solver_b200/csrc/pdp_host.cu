// Host-side ingest helpers of the C ABI (no device work): integer-list and DIMACS scanners for the
// predict path's input files.  They replace `json.loads` + `np.array(list)` (reference
// src/pdp/factorgraph/dataset.py:120-136) and the dense [m,n] clause matrix of src/dimacs2json.py:24-50,
// neither of which survives n = 1 M.
#include <cstdint>
#include <cstdlib>
#include "../../include/pdp_b200.h"

extern "C" {

int64_t pdp_host_parse_ints(const char* text, int64_t len, int32_t* out, int64_t cap) {
    if (!text || len < 0 || (cap > 0 && !out)) return -1;
    int64_t n = 0;
    int64_t i = 0;
    while (i < len) {
        unsigned d = (unsigned)(text[i] - '0');
        if (d > 9u) { ++i; continue; }
        bool neg = i > 0 && text[i - 1] == '-';
        int64_t v = 0;
        while (i < len && (d = (unsigned)(text[i] - '0')) <= 9u) { v = v * 10 + d; if (v > 0x7fffffffLL) return -2; ++i; }
        if (n < cap) out[n] = (int32_t)(neg ? -v : v);
        ++n;
    }
    return n;
}

int pdp_host_parse_dimacs(const char* text, int64_t len, int32_t* lits, int64_t cap, int64_t* info) {
    if (!text || len < 0 || !info || (cap > 0 && !lits)) return PDP_ERR_ARG;
    int64_t n = 0, clauses = 0, decl_v = -1, decl_c = -1;
    bool open_clause = false;
    int64_t i = 0;
    while (i < len) {
        // one line
        int64_t j = i;
        while (j < len && (text[j] == ' ' || text[j] == '\t' || text[j] == '\r')) ++j;
        int64_t eol = j;
        while (eol < len && text[eol] != '\n') ++eol;
        if (j < eol) {
            char c = text[j];
            if (c == '%') break;                                   // SATLIB end marker
            if (c == 'p') {
                int32_t hdr[2] = {0, 0};
                int64_t k = pdp_host_parse_ints(text + j, eol - j, hdr, 2);
                if (k < 2) return PDP_ERR_ARG;
                decl_v = hdr[0]; decl_c = hdr[1];
            } else if (c != 'c') {
                int64_t p = j;
                while (p < eol) {
                    unsigned d = (unsigned)(text[p] - '0');
                    if (d > 9u) { ++p; continue; }
                    bool neg = p > 0 && text[p - 1] == '-';
                    int64_t v = 0;
                    while (p < eol && (d = (unsigned)(text[p] - '0')) <= 9u) { v = v * 10 + d; if (v > 0x7fffffffLL) return PDP_ERR_ARG; ++p; }
                    if (n < cap) lits[n] = (int32_t)(neg ? -v : v);
                    ++n;
                    if (v == 0) { ++clauses; open_clause = false; } else open_clause = true;
                }
            }
        }
        i = eol + 1;
    }
    if (open_clause) {                                             // last clause without its terminator
        if (n < cap) lits[n] = 0;
        ++n; ++clauses;
    }
    info[0] = decl_v; info[1] = decl_c; info[2] = n; info[3] = clauses;
    return n <= cap ? PDP_OK : PDP_ERR_WORKSPACE;
}

// One pass over a whole compact-JSON file (one row per line):
//   [[n, m], [+-(variable+1) ...], [clause+1 ...], label, [id ...]]
// lits / cls receive the two integer lists of every row back to back; row_ptr[r] .. row_ptr[r+1] delimits row r in
// them; nm[2r], nm[2r+1] = n, m; label[r]; tail[2r], tail[2r+1] = byte range of what follows the label (the id list), for
// the caller to decode.  Returns the number of rows, or -(line number) of the first malformed row; -2^62 if a capacity is
// too small (rows > row_cap or integers > int_cap).
int64_t pdp_host_parse_rows(const char* text, int64_t len, int32_t* lits, int32_t* cls, int64_t int_cap,
                            int64_t* row_ptr, int32_t* nm, double* label, int64_t* tail, int64_t row_cap) {
    if (!text || len < 0 || !lits || !cls || !row_ptr || !nm || !label || !tail) return -1;
    const int64_t OVER = -((int64_t)1 << 62);
    int64_t rows = 0, nl = 0, nc = 0, line = 0;
    int64_t i = 0;
    row_ptr[0] = 0;
    while (i < len) {
        ++line;
        int64_t eol = i;
        while (eol < len && text[eol] != '\n') ++eol;
        int64_t p = i;
        while (p < eol && (text[p] == ' ' || text[p] == '\t' || text[p] == '\r')) ++p;
        if (p == eol) { i = eol + 1; continue; }                       // blank line
        if (rows >= row_cap) return OVER;
        // "[[n, m]"
        auto skip_to = [&](char ch) { while (p < eol && text[p] != ch) ++p; return p < eol; };
        auto read_int = [&](int64_t& v) {
            while (p < eol && text[p] != '-' && (unsigned)(text[p] - '0') > 9u && text[p] != ']') ++p;
            if (p >= eol || text[p] == ']') return false;
            bool neg = false;
            if (text[p] == '-') { neg = true; ++p; }
            if (p >= eol || (unsigned)(text[p] - '0') > 9u) return false;
            int64_t x = 0;
            while (p < eol && (unsigned)(text[p] - '0') <= 9u) { x = x * 10 + (text[p] - '0'); if (x > 0x7fffffffLL) return false; ++p; }
            v = neg ? -x : x;
            return true;
        };
        if (!skip_to('[')) return -line;
        ++p;
        if (!skip_to('[')) return -line;
        ++p;
        int64_t n = 0, m = 0;
        if (!read_int(n) || !read_int(m)) return -line;
        if (!skip_to(']')) return -line;
        ++p;
        // the two integer lists
        for (int which = 0; which < 2; ++which) {
            if (!skip_to('[')) return -line;
            ++p;
            int32_t* out = which == 0 ? lits : cls;
            int64_t k = which == 0 ? nl : nc;
            bool neg = false, closed = false;
            while (p < eol) {
                const unsigned d = (unsigned)(text[p] - '0');
                if (d <= 9u) {
                    int64_t x = d;
                    ++p;
                    unsigned e;
                    while (p < eol && (e = (unsigned)(text[p] - '0')) <= 9u) { x = x * 10 + e; ++p; }
                    if (x > 0x7fffffffLL) return -line;
                    if (k >= int_cap) return OVER;
                    out[k++] = (int32_t)(neg ? -x : x);
                    neg = false;
                } else {
                    const char ch = text[p++];
                    if (ch == '-') neg = true;
                    else if (ch == ']') { closed = true; break; }
                }
            }
            if (!closed) return -line;
            if (which == 0) nl = k; else nc = k;
        }
        if (nl != nc) return -line;
        // label
        while (p < eol && (text[p] == ',' || text[p] == ' ')) ++p;
        {
            char buf[64];
            int k = 0;
            while (p < eol && k < 63 && text[p] != ',' && text[p] != ']' && text[p] != ' ') buf[k++] = text[p++];
            buf[k] = 0;
            if (k == 0) return -line;
            char* endp = nullptr;
            label[rows] = strtod(buf, &endp);
            if (endp == buf) return -line;
        }
        while (p < eol && (text[p] == ',' || text[p] == ' ')) ++p;
        tail[2 * rows] = p; tail[2 * rows + 1] = eol;
        nm[2 * rows] = (int32_t)n; nm[2 * rows + 1] = (int32_t)m;
        ++rows;
        row_ptr[rows] = nl;
        i = eol + 1;
    }
    return rows;
}

}  // extern "C"
