/* compat shim (our own code): stands in for the cmake-generated basedescs.h (src/CMakeLists.txt:504-509).
 * Lists the reference filter descs that are compiled verbatim into oracle/_ref/libms2ref.so. */
#include "mediastreamer2/msfilter.h"
extern MSFilterDesc ms_void_source_desc;
extern MSFilterDesc ms_void_sink_desc;
extern MSFilterDesc ms_audio_mixer_desc;
extern MSFilterDesc ms_volume_desc;
extern MSFilterDesc ms_channel_adapter_desc;
extern MSFilterDesc ms_equalizer_desc;
extern MSFilterDesc ms_alaw_dec_desc, ms_alaw_enc_desc, ms_ulaw_dec_desc, ms_ulaw_enc_desc, ms_audio_flow_control_desc;
extern MSFilterDesc ms_genericplc_desc;
extern MSFilterDesc ms_pix_conv_desc, ms_size_conv_desc, ms_file_player_desc;
MSFilterDesc *ms_base_filter_descs[] = {&ms_void_source_desc, &ms_void_sink_desc, &ms_audio_mixer_desc,
                                        &ms_volume_desc, &ms_channel_adapter_desc, &ms_equalizer_desc,
                                        &ms_alaw_dec_desc, &ms_alaw_enc_desc, &ms_ulaw_dec_desc, &ms_ulaw_enc_desc,
                                        &ms_audio_flow_control_desc, &ms_genericplc_desc, &ms_pix_conv_desc,
                                        &ms_size_conv_desc, &ms_file_player_desc, NULL};
