/*
 * msb200_ms2.h — what a mediastreamer2 HOST needs to know about libmsb200filters.so beyond the reference's own headers:
 * the two pixel formats the reference's MSPixFmt lacks and the methods that make MSPixConv express them.
 *
 * The reference's MSPixConv converts camera formats to YUV420P at the same size (src/videofilters/pixconv.c:62-94; its
 * output format has no setter, :109-110) and MSPixFmt has no NV12 member (include/mediastreamer2/msvideo.h:267-280), so
 * BASELINE cfg4 (NV12 1080p -> RGB24 720p) cannot be asked of it. The plugin's MSPixConv keeps the reference's method
 * table and adds:
 */
#ifndef MSB200_MS2_H
#define MSB200_MS2_H

#include "mediastreamer2/allfilters.h"
#include "mediastreamer2/msfilter.h"
#include "mediastreamer2/msvideo.h"

/* appended after the last reference value (MS_RGBA32_REV), as SURVEY.md §8a proposes */
#define MSB200_MS_NV12 ((MSPixFmt)(MS_RGBA32_REV + 1)) /* Y plane, then interleaved CbCr */
#define MSB200_MS_NV21 ((MSPixFmt)(MS_RGBA32_REV + 2)) /* Y plane, then interleaved CrCb */

/* output format of MSPixConv: MS_YUV420P (default, the reference's only behaviour), MS_RGB24 or MS_RGB24_REV (BGR byte
 * order, rows top-down) — what the reference's display filters ask of ms_scaler_create_context (drawdib-display.c:76-103) */
#define MSB200_PIX_CONV_SET_OUTPUT_FMT MS_FILTER_METHOD(MS_PIX_CONV_ID, 64, MSPixFmt)
/* output size of MSPixConv: {0, 0} (default) = the input size; otherwise conversion and bilinear scaling are ONE kernel
 * (cfg4's NV12 1080p -> RGB24 720p reads 3.1 MB and writes 2.8 MB per frame instead of the two-filter chain's 10.7 MB) */
#define MSB200_PIX_CONV_SET_OUTPUT_SIZE MS_FILTER_METHOD(MS_PIX_CONV_ID, 65, MSVideoSize)

#endif
