/* compat shim (our own code) */
#define MS2_GIT_VERSION "compat"
