/* compat shim (our own code): the build-config header the reference generates with cmake.
 * Float build (no MS_FIXED_POINT), no speexdsp, no ffmpeg, no libyuv: only in-tree arithmetic is compiled. */
#ifndef MSB200_COMPAT_MS_CONFIG_H
#define MSB200_COMPAT_MS_CONFIG_H
#define MEDIASTREAMER_VERSION "5.5.0-compat"
#define NO_FFMPEG 1
#define PACKAGE_PLUGINS_DIR "/nonexistent/ms2plugins"
#define PACKAGE_DATA_DIR "/nonexistent"
#define HAVE_DLOPEN 1
#endif
