/* Test-infrastructure shim (NOT product code): minimal stand-in for libnuma's
 * <numa.h> so the unmodified reference src/numa_nodes.cc:8-11 compiles without
 * building the vendored numactl tarball (its autotools build is outside the
 * "few source files" rule).  Reports "NUMA unavailable", which the reference
 * handles itself (src/numa_nodes.cc:37-38,59-68): one node, bind() is a no-op.
 * Consequence: the reference arm always runs with -c 1 and all threads. */
#ifndef DWG_ORACLE_SHIM_NUMA_H_
#define DWG_ORACLE_SHIM_NUMA_H_
struct bitmask;
static inline int numa_available(void) { return -1; }
static inline int numa_num_configured_nodes(void) { return 1; }
static inline void numa_bind(struct bitmask *m) { (void)m; }
static inline struct bitmask *numa_parse_nodestring(const char *s) { (void)s; return 0; }
static inline void numa_bitmask_free(struct bitmask *m) { (void)m; }
#endif
