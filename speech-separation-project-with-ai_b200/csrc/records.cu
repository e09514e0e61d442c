// records.cu -- the feature / label record format the reference writes next to the path:
//   make_sequence_example(inputs, labels, length, name)      parallel_stft_single.py:238-254 (parallel_stft.py:217-229)
//   with tf.io.TFRecordWriter(path) as w: w.write(ex.SerializeToString())        :287-309
// i.e. one tf.train.SequenceExample whose feature_lists map holds
//   'inputs' : T features, each a FloatList of W_in  floats   (|X| ++ angle X, [T, 2F])
//   'labels' : T features, each a FloatList of W_lab floats   (PSA labels, [T, C F])
//   'length' : 1 feature, FloatList [length]                  (frames of the unpadded utterance)
//   'name'   : 1 feature, BytesList [utf-8 name]
// framed as a TFRecord: u64 length, u32 masked_crc32c(length), payload, u32 masked_crc32c(payload).
// Host code (byte formatting + CRC32C, no kernel): the arrays come from sep_stft_features_f32.  TensorFlow is not part
// of this stack; the encoder is checked byte for byte against the reference's committed .tfrecords files
// (tests/test_records.py).  Protobuf does not define the order of map entries -- the reference's own files differ in it
// from file to file -- so the caller may pass the order; the default is the sorted one.
#include <cstring>

#include "common.cuh"

namespace sep {

static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;

static void crc_init() {
  if (g_crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;   // CRC-32C (Castagnoli), reflected
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  g_crc_ready = true;
}

static uint32_t crc32c(const uint8_t *p, size_t n) {
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {                                               // slicing by 8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^
        g_crc_table[4][lo >> 24] ^ g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^
        g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_crc_table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

static uint32_t masked_crc(const uint8_t *p, size_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xA282EAD8u;                 // TFRecord's mask
}

static int varint_len(uint64_t v) {
  int n = 1;
  while (v >= 0x80) { v >>= 7; ++n; }
  return n;
}
static uint8_t *put_varint(uint8_t *p, uint64_t v) {
  while (v >= 0x80) { *p++ = static_cast<uint8_t>(v) | 0x80; v >>= 7; }
  *p++ = static_cast<uint8_t>(v);
  return p;
}
// length-delimited field: tag, varint(len) -- returns the header size
static int64_t ld_header(int64_t len) { return 1 + varint_len(static_cast<uint64_t>(len)); }

struct RecordLayout {
  int64_t float_body[2], feature[2], flist[4], entry[4], lists, payload;
};

static const char *kKeys[4] = {"inputs", "labels", "length", "name"};

static void layout(int frames, const int width[2], int name_len, RecordLayout *l) {
  for (int i = 0; i < 2; ++i) {
    const int64_t packed = 4LL * width[i];
    l->float_body[i] = width[i] > 0 ? ld_header(packed) + packed : 0;            // FloatList { value = 1 [packed] }
    l->feature[i] = ld_header(l->float_body[i]) + l->float_body[i];              // Feature { float_list = 2 }
    l->flist[i] = static_cast<int64_t>(frames) * (ld_header(l->feature[i]) + l->feature[i]);   // FeatureList { feature = 1 }
  }
  const int64_t len_feature = ld_header(6) + 6;                                  // FloatList of one float: 0A 04 ffff
  l->flist[2] = ld_header(len_feature) + len_feature;
  const int64_t bytes_body = ld_header(name_len) + name_len;                     // BytesList { value = 1 }
  const int64_t name_feature = ld_header(bytes_body) + bytes_body;               // Feature { bytes_list = 1 }
  l->flist[3] = ld_header(name_feature) + name_feature;
  l->lists = 0;
  for (int k = 0; k < 4; ++k) {
    const int64_t klen = static_cast<int64_t>(strlen(kKeys[k]));
    l->entry[k] = ld_header(klen) + klen + ld_header(l->flist[k]) + l->flist[k]; // map entry { key = 1, value = 2 }
    l->lists += ld_header(l->entry[k]) + l->entry[k];                            // FeatureLists { feature_list = 1 }
  }
  l->payload = ld_header(l->lists) + l->lists;                                   // SequenceExample { feature_lists = 2 }
}

static uint8_t *put_float_features(uint8_t *p, const float *data, int frames, int width, const RecordLayout &l, int which) {
  for (int t = 0; t < frames; ++t) {
    *p++ = 0x0A;                                                                 // FeatureList.feature
    p = put_varint(p, static_cast<uint64_t>(l.feature[which]));
    *p++ = 0x12;                                                                 // Feature.float_list
    p = put_varint(p, static_cast<uint64_t>(l.float_body[which]));
    if (width > 0) {
      *p++ = 0x0A;                                                               // FloatList.value, packed
      p = put_varint(p, 4ULL * width);
      memcpy(p, data + static_cast<int64_t>(t) * width, 4LL * width);            // little-endian host
      p += 4LL * width;
    }
  }
  return p;
}

}  // namespace sep

using namespace sep;

extern "C" uint32_t sep_record_masked_crc(const uint8_t *data, int64_t n) {
  return data && n >= 0 ? masked_crc(data, static_cast<size_t>(n)) : 0u;
}

extern "C" int sep_record_size(int frames, int width_inputs, int width_labels, int name_len, int64_t *bytes) {
  SEP_REQUIRE(bytes && frames >= 0 && width_inputs >= 0 && width_labels >= 0 && name_len >= 0,
              "sep_record_size: bad argument");
  RecordLayout l;
  const int width[2] = {width_inputs, width_labels};
  layout(frames, width, name_len, &l);
  *bytes = 8 + 4 + l.payload + 4;
  return SEP_OK;
}

extern "C" int sep_record_encode(const float *inputs, const float *labels, int frames, int width_inputs,
                                 int width_labels, float length, const char *name, int name_len,
                                 const int *key_order, uint8_t *out, int64_t capacity, int64_t *written) {
  SEP_REQUIRE(out && written && name && frames >= 0 && width_inputs >= 0 && width_labels >= 0 && name_len >= 0,
              "sep_record_encode: bad argument");
  SEP_REQUIRE((inputs || frames * width_inputs == 0) && (labels || frames * width_labels == 0),
              "sep_record_encode: null feature array");
  int order[4] = {0, 1, 2, 3};
  if (key_order) {
    int seen = 0;
    for (int k = 0; k < 4; ++k) {
      SEP_REQUIRE(key_order[k] >= 0 && key_order[k] < 4, "sep_record_encode: key_order must be a permutation of 0..3");
      seen |= 1 << key_order[k];
      order[k] = key_order[k];
    }
    SEP_REQUIRE(seen == 15, "sep_record_encode: key_order must be a permutation of 0..3");
  }
  RecordLayout l;
  const int width[2] = {width_inputs, width_labels};
  layout(frames, width, name_len, &l);
  const int64_t total = 8 + 4 + l.payload + 4;
  if (capacity < total) {
    set_error("sep_record_encode: buffer of %lld bytes, record needs %lld", static_cast<long long>(capacity),
              static_cast<long long>(total));
    return SEP_ERR_INVALID;
  }
  uint8_t *p = out;
  const uint64_t plen = static_cast<uint64_t>(l.payload);
  memcpy(p, &plen, 8);
  const uint32_t lcrc = masked_crc(p, 8);
  memcpy(p + 8, &lcrc, 4);
  p += 12;
  uint8_t *payload = p;
  *p++ = 0x12;                                                                   // SequenceExample.feature_lists
  p = put_varint(p, static_cast<uint64_t>(l.lists));
  for (int q = 0; q < 4; ++q) {
    const int k = order[q];
    const size_t klen = strlen(kKeys[k]);
    *p++ = 0x0A;                                                                 // FeatureLists.feature_list (map entry)
    p = put_varint(p, static_cast<uint64_t>(l.entry[k]));
    *p++ = 0x0A;                                                                 // key
    p = put_varint(p, klen);
    memcpy(p, kKeys[k], klen);
    p += klen;
    *p++ = 0x12;                                                                 // value: FeatureList
    p = put_varint(p, static_cast<uint64_t>(l.flist[k]));
    if (k == 0) p = put_float_features(p, inputs, frames, width_inputs, l, 0);
    else if (k == 1) p = put_float_features(p, labels, frames, width_labels, l, 1);
    else if (k == 2) {
      const uint8_t head[6] = {0x0A, 0x08, 0x12, 0x06, 0x0A, 0x04};              // feature(8) { float_list(6) { value(4) } }
      memcpy(p, head, 6);
      memcpy(p + 6, &length, 4);
      p += 10;
    } else {
      const int64_t bytes_body = ld_header(name_len) + name_len;
      *p++ = 0x0A;                                                               // FeatureList.feature
      p = put_varint(p, static_cast<uint64_t>(ld_header(bytes_body) + bytes_body));
      *p++ = 0x0A;                                                               // Feature.bytes_list
      p = put_varint(p, static_cast<uint64_t>(bytes_body));
      *p++ = 0x0A;                                                               // BytesList.value
      p = put_varint(p, static_cast<uint64_t>(name_len));
      memcpy(p, name, name_len);
      p += name_len;
    }
  }
  if (p - payload != l.payload) {
    set_error("sep_record_encode: internal size mismatch (%lld vs %lld)", static_cast<long long>(p - payload),
              static_cast<long long>(l.payload));
    return SEP_ERR_INVALID;
  }
  const uint32_t pcrc = masked_crc(payload, static_cast<size_t>(l.payload));
  memcpy(p, &pcrc, 4);
  *written = total;
  return SEP_OK;
}
