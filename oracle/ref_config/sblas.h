/* empty stand-in: NEC VE-only header included unconditionally by the reference (src/seq_mv/csr_matvec.c:16-17, src/parcsr_ls/par_relax.c) */
