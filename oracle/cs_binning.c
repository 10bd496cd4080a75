/*
 * cs_binning.c — CPU restatement of gene -> bin count aggregation.
 *
 * TEST INFRASTRUCTURE ONLY (see cs_oracle.h).
 *
 * Reference: R/Gene_bin_cell_rna_filtered.R:63-81.  `ov = findOverlaps(query,
 * subject)` lists (bin, gene) pairs (any overlap, many-to-many); row ii of the
 * result is the column sum of the genes that overlap bin ii (:72-77); the dense
 * result is then converted with Matrix(., sparse=TRUE) (:81), i.e. a dgCMatrix
 * with ascending row indices per column and no stored zeros.
 *
 * The overlap table itself is integer interval logic done on the host
 * (clonalscope_b200/intervals.py mirrors :30-64); here it arrives as a CSR map
 * gene -> list of bins (map_ptr/map_bin).  Parity of this stage is NOT pinned
 * by reference data (every count matrix is in .MISSING_LARGE_BLOBS); the two
 * restatements below (sparse sweep and the reference's own dense loop) are
 * checked against each other.
 */
#include "cs_oracle.h"

#include <stdlib.h>
#include <string.h>

int64_t cso_bin_counts(int G, int N, int B, const int64_t *colptr, const int32_t *rowidx,
                       const double *x, const int32_t *map_ptr, const int32_t *map_bin,
                       int64_t *out_colptr, int32_t *out_rowidx, double *out_x)
{
    (void)G;
    long double *acc = (long double *)calloc((size_t)B, sizeof(long double));
    unsigned char *touched = (unsigned char *)calloc((size_t)B, 1);
    int32_t *list = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    int64_t nnz = 0;
    out_colptr[0] = 0;
    for (int c = 0; c < N; c++) {
        int nt = 0;
        for (int64_t e = colptr[c]; e < colptr[c + 1]; e++) {
            int g = rowidx[e];
            for (int32_t m = map_ptr[g]; m < map_ptr[g + 1]; m++) {
                int b = map_bin[m];
                if (!touched[b]) { touched[b] = 1; list[nt++] = b; }
                acc[b] += x[e]; /* ascending gene order within each bin */
            }
        }
        /* emit in ascending bin order, dropping zero sums like Matrix(sparse=TRUE) */
        /* small insertion-friendly sort: lists are nearly sorted */
        for (int a = 1; a < nt; a++) {
            int32_t v = list[a];
            int j = a - 1;
            while (j >= 0 && list[j] > v) { list[j + 1] = list[j]; j--; }
            list[j + 1] = v;
        }
        for (int a = 0; a < nt; a++) {
            int b = list[a];
            double v = (double)acc[b];
            acc[b] = 0.0L;
            touched[b] = 0;
            if (v != 0.0) {
                if (out_rowidx) { out_rowidx[nnz] = b; out_x[nnz] = v; }
                nnz++;
            }
        }
        out_colptr[c + 1] = nnz;
    }
    free(acc); free(touched); free(list);
    return nnz;
}

void cso_bin_counts_dense(int G, int N, int B, const int64_t *colptr, const int32_t *rowidx,
                          const double *x, const int32_t *map_ptr, const int32_t *map_bin,
                          double *out)
{
    /* :69  mtx = as.matrix(gene_matrix)  (dense genes x cells, column-major) */
    double *dense = (double *)calloc((size_t)G * N, sizeof(double));
    for (int c = 0; c < N; c++)
        for (int64_t e = colptr[c]; e < colptr[c + 1]; e++)
            dense[(size_t)c * G + rowidx[e]] += x[e];
    /* invert the gene->bin map into bin -> ascending gene list (= ov[ov[,1]==ii, 2]) */
    int32_t *bptr = (int32_t *)calloc((size_t)B + 1, sizeof(int32_t));
    for (int g = 0; g < G; g++)
        for (int32_t m = map_ptr[g]; m < map_ptr[g + 1]; m++) bptr[map_bin[m] + 1]++;
    for (int b = 0; b < B; b++) bptr[b + 1] += bptr[b];
    int32_t *bgene = (int32_t *)malloc(sizeof(int32_t) * (size_t)(bptr[B] > 0 ? bptr[B] : 1));
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)B);
    memcpy(fill, bptr, sizeof(int32_t) * (size_t)B);
    for (int g = 0; g < G; g++)
        for (int32_t m = map_ptr[g]; m < map_ptr[g + 1]; m++) bgene[fill[map_bin[m]]++] = g;
    /* :72-77  one bin at a time, column sums over its genes */
    for (int b = 0; b < B; b++)
        for (int c = 0; c < N; c++) {
            long double s = 0.0L;
            for (int32_t m = bptr[b]; m < bptr[b + 1]; m++) s += dense[(size_t)c * G + bgene[m]];
            out[(size_t)c * B + b] = (double)s;
        }
    free(dense); free(bptr); free(bgene); free(fill);
}
