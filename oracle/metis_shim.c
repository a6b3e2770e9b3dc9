/* ORACLE-SIDE METIS access (test infrastructure): re-exports the bundled static METIS so the
 * tests can call it with the reference's arguments (ExodusIO.hpp:1615). */
int oracle_metis_shim_present(void) { return 1; }
