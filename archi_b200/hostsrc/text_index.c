/*
 * text_index.c -- host-side helper of the lexical (BM25) index: tokenise a batch of chunks and
 * reduce every chunk to its distinct terms with their frequencies.  Plain C, no CUDA: it feeds
 * archi_bm25_accumulate's posting lists (archi_b200/bm25.py builds the CSR from its output).
 *
 * Replaces, on the ingest side, what pg_textsearch does when a row is inserted into a table with a
 * `USING bm25` index (src/cli/templates/init.sql:297-300) [external, parity unpinned: tokenisation
 * here is lower-cased alphanumeric runs, see DESIGN.md assumptions].
 *
 * A term is identified by the 64-bit FNV-1a hash of its lower-cased bytes (collision odds for a
 * vocabulary of 10^6 terms: ~3e-8), so no dictionary is touched while indexing.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* class of a byte: 0 = separator, otherwise the lower-cased byte */
static unsigned char g_lower[256];
static int g_ready = 0;
static void init_table(void)
{
    for (int c = 0; c < 256; c++) {
        if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z')) g_lower[c] = (unsigned char)c;
        else if (c >= 'A' && c <= 'Z') g_lower[c] = (unsigned char)(c + 32);
        else g_lower[c] = 0;               /* everything else, bytes >= 0x80 included, separates */
    }
    g_ready = 1;
}

/*
 * text            concatenated bytes of n_docs documents (ASCII; other bytes act as separators)
 * offs[n_docs+1]  byte offsets of the documents inside text
 * keys_out / tfs_out [cap]   distinct term keys of each document, ascending, and their counts
 * doc_ptr_out[n_docs+1]      where each document's pairs start in keys_out / tfs_out
 * doc_len_out[n_docs]        number of tokens of each document
 * Returns the number of pairs written, -1 if cap is too small, -2 on allocation failure.
 */
int64_t archi_text_index_batch(const uint8_t *text, const int64_t *offs, int64_t n_docs, uint64_t *keys_out,
                               int32_t *tfs_out, int64_t cap, int64_t *doc_ptr_out, int32_t *doc_len_out)
{
    if (!g_ready) init_table();
    int64_t scratch_cap = 1024;
    uint64_t *scratch = (uint64_t *)malloc((size_t)scratch_cap * sizeof(uint64_t));
    if (!scratch) return -2;
    int64_t w = 0;
    for (int64_t d = 0; d < n_docs; d++) {
        const uint8_t *p = text + offs[d], *end = text + offs[d + 1];
        int64_t nt = 0;
        while (p < end) {
            while (p < end && !g_lower[*p]) p++;
            if (p >= end) break;
            uint64_t h = 1469598103934665603ull;             /* FNV-1a 64 */
            while (p < end && g_lower[*p]) {
                h ^= g_lower[*p];
                h *= 1099511628211ull;
                p++;
            }
            if (nt == scratch_cap) {
                scratch_cap *= 2;
                uint64_t *grown = (uint64_t *)realloc(scratch, (size_t)scratch_cap * sizeof(uint64_t));
                if (!grown) {
                    free(scratch);
                    return -2;
                }
                scratch = grown;
            }
            scratch[nt++] = h;
        }
        doc_ptr_out[d] = w;
        doc_len_out[d] = (int32_t)nt;
        if (nt > 1) qsort(scratch, (size_t)nt, sizeof(uint64_t), cmp_u64);
        for (int64_t i = 0; i < nt;) {
            int64_t j = i + 1;
            while (j < nt && scratch[j] == scratch[i]) j++;
            if (w >= cap) {
                free(scratch);
                return -1;
            }
            keys_out[w] = scratch[i];
            tfs_out[w] = (int32_t)(j - i);
            w++;
            i = j;
        }
    }
    doc_ptr_out[n_docs] = w;
    free(scratch);
    return w;
}

/* The key of one already lower-cased alphanumeric token (used for query terms and by the tests). */
uint64_t archi_text_term_key(const uint8_t *token, int64_t len)
{
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < len; i++) {
        h ^= token[i];
        h *= 1099511628211ull;
    }
    return h;
}
