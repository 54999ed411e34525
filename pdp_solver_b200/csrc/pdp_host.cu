// Host-side ingest helpers of the C ABI (no device work): integer-list and DIMACS scanners for the
// predict path's input files.  They replace `json.loads` + `np.array(list)` (reference
// src/pdp/factorgraph/dataset.py:120-136) and the dense [m,n] clause matrix of src/dimacs2json.py:24-50,
// neither of which survives n = 1 M.
#include <cstdint>
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

}  // extern "C"
