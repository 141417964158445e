/* plugin/msb200_plugin.h — shared between the translation units of libmsb200filters.so (not installed) */
#ifndef MSB200_PLUGIN_H
#define MSB200_PLUGIN_H

#include "mediastreamer2/msfactory.h"
#include "mediastreamer2/msfilter.h"
#include "mediastreamer2/msticker.h"
#include "mediastreamer2/msvideo.h"

#include "msb200dsp.h"

/* the process-wide context of the synchronous filters (MSB200_DEVICE); NULL when no GPU. Take the lock around every
 * bank call made on it: its stream is shared by all synchronous filter instances. */
msb200_ctx *msb200p_sync_ctx(void);
void msb200p_sync_lock(void);
void msb200p_sync_unlock(void);
/* MSB200_BATCH=<slots>: > 0 turns the lockstep batch mode on */
int msb200p_batch_capacity(void);
/* MSB200_DEVICES=<n>: the device a ticker's batch groups live on */
int msb200p_device_of_ticker(MSTicker *t);

/* msb200_video_filters.c */
void msb200p_register_video_filters(MSFactory *factory);
MSScalerDesc *msb200p_scaler_desc(void);
int msb200p_pixfmt_to_b200(MSPixFmt fmt);

#endif
