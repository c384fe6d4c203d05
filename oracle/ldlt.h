/* ldlt.h -- internal interface of the oracle's sparse LDL^T (TEST INFRASTRUCTURE ONLY). */
#ifndef SIM3OPT_ORACLE_LDLT_H
#define SIM3OPT_ORACLE_LDLT_H
typedef struct orc_ldlt orc_ldlt;
/* nb x nb block matrix of d x d blocks, upper block-CCS pattern (rows <= col) */
orc_ldlt *orc_ldlt_analyze(int nb, int d, const int *colptr, const int *rowidx);
/* blocks: colptr[nb] blocks in CCS order, each row-major d x d; factors (A + lambda I) */
int orc_ldlt_factor(orc_ldlt *S, const double *blocks, double lambda);
void orc_ldlt_solve(orc_ldlt *S, const double *b, double *x);
long orc_ldlt_lnz(const orc_ldlt *S);
void orc_ldlt_free(orc_ldlt *S);
#endif
