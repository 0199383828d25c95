"""B200-native cloud_codec_v2 intra encode/decode hot path."""
