/* bpl_oracle.h -- TEST INFRASTRUCTURE (CPU oracle for the pyBPL render+score path).
 * Constants mirror pybpl/parameters.py:19-41 of the reference. */
#ifndef BPL_ORACLE_H
#define BPL_ORACLE_H

typedef struct {
    int H, W;            /* imsize (parameters.py:21) */
    float ink_pp;        /* :25 */
    float ink_max_dist;  /* :27 */
    int ink_ncon;        /* :31 */
    float ink_a, ink_b;  /* :33,:35 */
    int hinton;          /* broaden_mode == 'Hinton' (:37) */
    int fsize;           /* :41 */
} bplo_params;

int bplo_version(void);
void bplo_default_params(bplo_params *ps);
int bplo_set_threads(int n);

#endif
