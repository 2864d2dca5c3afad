/* Test-infrastructure shim: empty stand-in for libnuma's <numaif.h>
 * (included by the reference's src/numa_nodes.cc:10, nothing from it is used). */
