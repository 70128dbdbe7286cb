// Dependency-free ROS1 bag (format "#ROSBAG V2.0") reader + minimal writer, and the three message types the LINS front
// end consumes.  PRODUCT code: what lets BASELINE.json configs[1] (one real scan of
// lidar_imu_dataset.bag, /root/reference/README.md:51) run the moment the bag is supplied — no ROS needed.
//
// The reference reads these through ROS subscribers (lins/src/lib/Estimator.cpp:36-45: IMU_TOPIC sensor_msgs/Imu,
// "/segmented_cloud" + "/outlier_cloud" sensor_msgs/PointCloud2, "/segmented_cloud_info" cloud_msgs/cloud_info;
// lins/src/image_projection_node.cpp:83: LIDAR_TOPIC sensor_msgs/PointCloud2) and converts clouds with pcl::fromROSMsg.
//
// Bag 2.0 layout (public format description, wiki.ros.org/Bags/Format/2.0): the magic line, then records; a record is
//   <u32 header_len> <header> <u32 data_len> <data>,  header = fields  <u32 field_len> name '=' value.
// Record kinds by the 1-byte field "op": 0x03 bag header (index_pos, conn_count, chunk_count; padded to 4096 bytes),
// 0x05 chunk (compression, size; data = connection + message-data records), 0x07 connection (conn, topic; data =
// connection header with type / md5sum / message_definition), 0x02 message data (conn, time; data = serialised message),
// 0x04 index data, 0x06 chunk info.  All integers little endian; time = u32 sec + u32 nsec.
// This reader walks the records in file order (the index sections are skipped, so truncated / unindexed bags still
// read); chunks with compression "none" are parsed in place, "lz4" chunks (one LZ4 frame per chunk, as ros_comm's roslz4
// writes them) are inflated by the decoder below, "bz2" chunks are reported as LINS_BAG_E_COMPRESSED
// (`tools/bag_tool.py decompress` rewrites such a bag with Python's bz2 module).
#ifndef LINS_HOST_ROSBAG_READER_HPP_
#define LINS_HOST_ROSBAG_READER_HPP_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "cloud.hpp"

namespace lins {
namespace rosbag {

enum { LINS_BAG_OK = 0, LINS_BAG_E_IO = -1, LINS_BAG_E_FORMAT = -2, LINS_BAG_E_COMPRESSED = -3 };

struct Connection {
  uint32_t id = 0;
  std::string topic, type, md5sum;
};

struct MessageView {
  const Connection* conn = nullptr;
  double time = 0;  // record time (when the message was recorded)
  const uint8_t* data = nullptr;
  size_t size = 0;
};

// ---- byte-level helpers -------------------------------------------------------------------------------------------------
struct Cursor {
  const uint8_t* p;
  const uint8_t* e;
  bool ok = true;
  Cursor(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  size_t left() const { return (size_t)(e - p); }
  template <typename T>
  T get() {
    T v{};
    if (left() < sizeof(T)) { ok = false; p = e; return v; }
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const uint8_t* bytes(size_t n) {
    if (left() < n) { ok = false; p = e; return nullptr; }
    const uint8_t* r = p;
    p += n;
    return r;
  }
  std::string str() {  // ROS string: u32 length + bytes
    const uint32_t n = get<uint32_t>();
    const uint8_t* b = bytes(n);
    return b ? std::string(reinterpret_cast<const char*>(b), n) : std::string();
  }
  template <typename T>
  std::vector<T> array() {  // ROS variable-length array of a fixed-size primitive
    const uint32_t n = get<uint32_t>();
    std::vector<T> v;
    if ((size_t)n * sizeof(T) > left()) { ok = false; p = e; return v; }
    v.resize(n);
    if (n) std::memcpy(v.data(), p, (size_t)n * sizeof(T));
    p += (size_t)n * sizeof(T);
    return v;
  }
};

typedef std::map<std::string, std::string> Fields;
inline bool parse_fields(const uint8_t* b, size_t n, Fields& out) {
  Cursor c(b, n);
  while (c.left() > 0) {
    const uint32_t fl = c.get<uint32_t>();
    const uint8_t* f = c.bytes(fl);
    if (!f) return false;
    const void* eq = std::memchr(f, '=', fl);
    if (!eq) return false;
    const size_t nl = (size_t)(static_cast<const uint8_t*>(eq) - f);
    out[std::string(reinterpret_cast<const char*>(f), nl)] = std::string(reinterpret_cast<const char*>(f) + nl + 1, fl - nl - 1);
  }
  return c.ok;
}
template <typename T>
inline bool field_as(const Fields& f, const char* name, T& v) {
  auto it = f.find(name);
  if (it == f.end() || it->second.size() < sizeof(T)) return false;
  std::memcpy(&v, it->second.data(), sizeof(T));
  return true;
}

// ---- LZ4 (public formats) ----------------------------------------------------------------------------------------------
// Block: sequences of <token: literal length << 4 | match length - 4> [length bytes while 255] <literals> <u16 offset>
// [length bytes]; a match copies byte by byte from `offset` back in the output (it may overlap itself); the last sequence
// has literals only.  Appends to `out`, never beyond `limit` bytes; earlier output stays addressable (linked blocks).
inline bool lz4_block_decode(const uint8_t* s, size_t n, std::vector<uint8_t>& out, size_t limit) {
  const uint8_t* const e = s + n;
  while (s < e) {
    const unsigned tok = *s++;
    size_t ll = tok >> 4;
    if (ll == 15) {
      unsigned b;
      do { if (s >= e) return false; b = *s++; ll += b; } while (b == 255);
    }
    if ((size_t)(e - s) < ll || limit - out.size() < ll) return false;
    out.insert(out.end(), s, s + ll);
    s += ll;
    if (s >= e) break;
    if (e - s < 2) return false;
    const size_t off = (size_t)s[0] | ((size_t)s[1] << 8);
    s += 2;
    if (off == 0 || off > out.size()) return false;
    size_t ml = tok & 15u;
    if (ml == 15) {
      unsigned b;
      do { if (s >= e) return false; b = *s++; ml += b; } while (b == 255);
    }
    ml += 4;
    if (limit - out.size() < ml) return false;
    const size_t from = out.size() - off;
    out.resize(out.size() + ml);
    uint8_t* d = out.data() + from + off;
    const uint8_t* m = out.data() + from;
    for (size_t i = 0; i < ml; ++i) d[i] = m[i];
  }
  return true;
}
// Frame: magic 0x184D2204, FLG (version 01, block independence, block checksum, content size, content checksum, dict id),
// BD, [u64 content size], [u32 dict id], header checksum, blocks <u32 size; bit 31 = stored raw> [u32 block checksum],
// end mark 0, [u32 content checksum].  Checksums are skipped, not verified (the chunk's record carries its inflated size).
inline bool lz4_frame_decode(const uint8_t* s, size_t n, std::vector<uint8_t>& out, size_t limit) {
  Cursor c(s, n);
  if (c.get<uint32_t>() != 0x184D2204u) return false;
  const uint8_t flg = c.get<uint8_t>();
  c.get<uint8_t>();  // BD: the block maximum size (the output limit bounds the work instead)
  if (!c.ok || (flg >> 6) != 1) return false;
  if (flg & 0x08) c.get<uint64_t>();
  if (flg & 0x01) c.get<uint32_t>();
  c.get<uint8_t>();
  for (;;) {
    uint32_t bs = c.get<uint32_t>();
    if (!c.ok) return false;
    if (bs == 0) return true;
    const bool raw = (bs >> 31) != 0;
    bs &= 0x7fffffffu;
    const uint8_t* b = c.bytes(bs);
    if (!b) return false;
    if (raw) {
      if (limit - out.size() < bs) return false;
      out.insert(out.end(), b, b + bs);
    } else if (!lz4_block_decode(b, bs, out, limit)) {
      return false;
    }
    if (flg & 0x10) c.get<uint32_t>();
  }
}

// ---- reader -------------------------------------------------------------------------------------------------------------
class Reader {
 public:
  std::map<uint32_t, Connection> connections;
  std::string error;
  uint32_t declared_conn_count = 0, declared_chunk_count = 0;

  int open(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { error = "cannot open " + path; return LINS_BAG_E_IO; }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf_.resize(n > 0 ? (size_t)n : 0);
    const size_t got = buf_.empty() ? 0 : std::fread(buf_.data(), 1, buf_.size(), f);
    std::fclose(f);
    if (got != buf_.size()) { error = "short read"; return LINS_BAG_E_IO; }
    static const char magic[] = "#ROSBAG V2.0\n";
    if (buf_.size() < sizeof(magic) - 1 || std::memcmp(buf_.data(), magic, sizeof(magic) - 1) != 0) { error = "not a ROS bag v2.0"; return LINS_BAG_E_FORMAT; }
    start_ = sizeof(magic) - 1;
    return LINS_BAG_OK;
  }

  // Calls fn for every message in file order (within a chunk: record order = recording order).
  int for_each(const std::function<void(const MessageView&)>& fn) {
    return walk(buf_.data() + start_, buf_.size() - start_, fn, true);
  }

 private:
  std::vector<uint8_t> buf_;
  size_t start_ = 0;

  int walk(const uint8_t* b, size_t n, const std::function<void(const MessageView&)>& fn, bool top) {
    Cursor c(b, n);
    while (c.left() >= 4) {
      const uint32_t hl = c.get<uint32_t>();
      const uint8_t* hb = c.bytes(hl);
      const uint32_t dl = c.get<uint32_t>();
      const uint8_t* db = c.bytes(dl);
      if (!c.ok || !hb || (dl && !db)) { error = "truncated record"; return top ? LINS_BAG_OK : LINS_BAG_E_FORMAT; }  // a bag cut short still yields its complete records
      Fields f;
      if (!parse_fields(hb, hl, f)) { error = "bad record header"; return LINS_BAG_E_FORMAT; }
      uint8_t op = 0;
      if (!field_as(f, "op", op)) { error = "record without op"; return LINS_BAG_E_FORMAT; }
      switch (op) {
        case 0x03:
          field_as(f, "conn_count", declared_conn_count);
          field_as(f, "chunk_count", declared_chunk_count);
          break;
        case 0x05: {
          const std::string comp = f.count("compression") ? f["compression"] : "none";
          if (comp == "lz4") {
            uint32_t size = 0;
            if (!field_as(f, "size", size) || size > (1u << 30)) { error = "lz4 chunk without a sane size"; return LINS_BAG_E_FORMAT; }
            std::vector<uint8_t> inflated;  // (messages are handed to fn during the walk: the buffer only has to outlive it)
            inflated.reserve(size);
            if (!lz4_frame_decode(db, dl, inflated, size) || inflated.size() != size) { error = "corrupt lz4 chunk"; return LINS_BAG_E_FORMAT; }
            const int rc = walk(inflated.data(), inflated.size(), fn, false);
            if (rc != LINS_BAG_OK) return rc;
            break;
          }
          if (comp != "none") { error = "chunk compression '" + comp + "' (run tools/bag_tool.py decompress first)"; return LINS_BAG_E_COMPRESSED; }
          const int rc = walk(db, dl, fn, false);
          if (rc != LINS_BAG_OK) return rc;
          break;
        }
        case 0x07: {
          Connection cn;
          field_as(f, "conn", cn.id);
          cn.topic = f.count("topic") ? f["topic"] : "";
          Fields ch;
          if (!parse_fields(db, dl, ch)) { error = "bad connection header"; return LINS_BAG_E_FORMAT; }
          cn.type = ch["type"]; cn.md5sum = ch["md5sum"];
          if (ch.count("topic") && cn.topic.empty()) cn.topic = ch["topic"];
          connections[cn.id] = cn;
          break;
        }
        case 0x02: {
          uint32_t id = 0;
          uint64_t t = 0;
          field_as(f, "conn", id);
          field_as(f, "time", t);
          auto it = connections.find(id);
          if (it == connections.end()) break;  // message of an undeclared connection: skip
          MessageView m;
          m.conn = &it->second;
          m.time = (double)(uint32_t)(t & 0xffffffffu) + 1e-9 * (double)(uint32_t)(t >> 32);
          m.data = db; m.size = dl;
          fn(m);
          break;
        }
        default: break;  // 0x04 index data, 0x06 chunk info: not needed for a sequential read
      }
    }
    return LINS_BAG_OK;
  }
};

// ---- messages -----------------------------------------------------------------------------------------------------------
struct Header {
  uint32_t seq = 0;
  double stamp = 0;
  std::string frame_id;
};
inline Header read_header(Cursor& c) {
  Header h;
  h.seq = c.get<uint32_t>();
  const uint32_t s = c.get<uint32_t>(), ns = c.get<uint32_t>();
  h.stamp = (double)s + 1e-9 * (double)ns;
  h.frame_id = c.str();
  return h;
}

// sensor_msgs/Imu -> what LinsFusion::imuCallback reads (Estimator.cpp:123-131)
struct ImuMsg {
  Header header;
  double orientation[4];  // x y z w
  double angular_velocity[3], linear_acceleration[3];
};
inline bool decode_imu(const uint8_t* b, size_t n, ImuMsg& m) {
  Cursor c(b, n);
  m.header = read_header(c);
  for (double& v : m.orientation) v = c.get<double>();
  c.bytes(9 * 8);
  for (double& v : m.angular_velocity) v = c.get<double>();
  c.bytes(9 * 8);
  for (double& v : m.linear_acceleration) v = c.get<double>();
  c.bytes(9 * 8);
  return c.ok;
}

// sensor_msgs/PointCloud2 -> PointXYZI cloud, as pcl::fromROSMsg<pcl::PointXYZI> maps fields by name (x, y, z,
// intensity; other fields such as the Velodyne driver's "ring" are ignored; a missing intensity stays 0).
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0; };
inline double read_scalar(const uint8_t* p, uint8_t datatype) {
  switch (datatype) {  // sensor_msgs/PointField constants
    case 1: { int8_t v; std::memcpy(&v, p, 1); return v; }
    case 2: { uint8_t v; std::memcpy(&v, p, 1); return v; }
    case 3: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case 4: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case 5: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case 6: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case 7: { float v; std::memcpy(&v, p, 4); return v; }
    case 8: { double v; std::memcpy(&v, p, 8); return v; }
  }
  return 0.0;
}
inline bool decode_pointcloud2(const uint8_t* b, size_t n, Header& hdr, Cloud& out, bool* is_dense = nullptr) {
  Cursor c(b, n);
  hdr = read_header(c);
  const uint32_t height = c.get<uint32_t>(), width = c.get<uint32_t>();
  const uint32_t nf = c.get<uint32_t>();
  if (!c.ok || nf > 64) return false;
  std::vector<PointField> fields(nf);
  for (auto& f : fields) { f.name = c.str(); f.offset = c.get<uint32_t>(); f.datatype = c.get<uint8_t>(); f.count = c.get<uint32_t>(); }
  const uint8_t bigendian = c.get<uint8_t>();
  const uint32_t point_step = c.get<uint32_t>(), row_step = c.get<uint32_t>();
  const uint32_t dlen = c.get<uint32_t>();
  const uint8_t* d = c.bytes(dlen);
  const uint8_t dense = c.get<uint8_t>();
  if (!c.ok || bigendian || (dlen && !d)) return false;
  if (is_dense) *is_dense = dense != 0;
  const PointField *fx = nullptr, *fy = nullptr, *fz = nullptr, *fi = nullptr;
  for (const auto& f : fields) {
    if (f.name == "x") fx = &f; else if (f.name == "y") fy = &f; else if (f.name == "z") fz = &f; else if (f.name == "intensity") fi = &f;
  }
  if (!fx || !fy || !fz) return false;
  static const uint32_t tsize[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
  for (const PointField* f : {fx, fy, fz, fi})
    if (f && (f->datatype < 1 || f->datatype > 8 || f->offset + tsize[f->datatype] > point_step)) return false;
  out.clear();
  out.points.reserve((size_t)width * height);
  for (uint32_t r = 0; r < height; ++r)
    for (uint32_t col = 0; col < width; ++col) {
      const size_t o = (size_t)r * row_step + (size_t)col * point_step;
      if (o + point_step > dlen) return false;
      const uint8_t* p = d + o;
      out.push_back(makePoint((float)read_scalar(p + fx->offset, fx->datatype), (float)read_scalar(p + fy->offset, fy->datatype),
                              (float)read_scalar(p + fz->offset, fz->datatype), fi ? (float)read_scalar(p + fi->offset, fi->datatype) : 0.f));
    }
  return true;
}

// cloud_msgs/cloud_info (cloud_msgs/msg/cloud_info.msg:1-12)
inline bool decode_cloud_info(const uint8_t* b, size_t n, Header& hdr, CloudInfo& ci) {
  Cursor c(b, n);
  hdr = read_header(c);
  ci.startRingIndex = c.array<int32_t>();
  ci.endRingIndex = c.array<int32_t>();
  ci.startOrientation = c.get<float>();
  ci.endOrientation = c.get<float>();
  ci.orientationDiff = c.get<float>();
  ci.segmentedCloudGroundFlag = c.array<uint8_t>();
  ci.segmentedCloudColInd = c.array<uint32_t>();
  ci.segmentedCloudRange = c.array<float>();
  return c.ok;
}

// ---- minimal writer (uncompressed, one chunk per flush, with index + chunk-info sections so that rosbag tools accept it) ---
class Writer {
 public:
  int open(const std::string& path) {
    f_ = std::fopen(path.c_str(), "wb");
    if (!f_) return LINS_BAG_E_IO;
    std::fputs("#ROSBAG V2.0\n", f_);
    write_bag_header(0, 0, 0);  // placeholder, rewritten by close()
    return LINS_BAG_OK;
  }
  uint32_t add_connection(const std::string& topic, const std::string& type, const std::string& md5, const std::string& definition) {
    const uint32_t id = (uint32_t)conns_.size();
    conns_.push_back({id, topic, type, md5, definition});
    return id;
  }
  void write(uint32_t conn, double time, const std::vector<uint8_t>& msg) {
    if (!announced_.count(conn)) { append_connection(chunk_, conns_[conn]); announced_[conn] = true; }
    const uint64_t t = to_time(time);
    index_[conn].push_back({t, (uint32_t)chunk_.size()});
    std::vector<uint8_t> h;
    put_field(h, "op", std::string(1, (char)0x02));
    put_field(h, "conn", raw(conn));
    put_field(h, "time", raw(t));
    put_record(chunk_, h, msg);
    if (t < t0_) t0_ = t;
    if (t > t1_) t1_ = t;
    if (chunk_.size() > (768u << 10)) flush_chunk();
  }
  int close() {
    if (!f_) return LINS_BAG_E_IO;
    flush_chunk();
    const uint64_t index_pos = (uint64_t)std::ftell(f_);
    std::vector<uint8_t> tail;
    for (const auto& c : conns_) append_connection(tail, c);
    for (const auto& ci : chunk_infos_) {
      std::vector<uint8_t> h, d;
      put_field(h, "op", std::string(1, (char)0x06));
      put_field(h, "ver", raw((uint32_t)1));
      put_field(h, "chunk_pos", raw(ci.pos));
      put_field(h, "start_time", raw(ci.t0));
      put_field(h, "end_time", raw(ci.t1));
      put_field(h, "count", raw((uint32_t)ci.counts.size()));
      for (const auto& kv : ci.counts) { append(d, raw(kv.first)); append(d, raw(kv.second)); }
      put_record(tail, h, d);
    }
    std::fwrite(tail.data(), 1, tail.size(), f_);
    std::fseek(f_, 13, SEEK_SET);
    write_bag_header(index_pos, (uint32_t)conns_.size(), (uint32_t)chunk_infos_.size());
    std::fclose(f_);
    f_ = nullptr;
    return LINS_BAG_OK;
  }

  // serialisers of the three message types (inverse of the decoders above)
  static void put_header(std::vector<uint8_t>& o, uint32_t seq, double stamp, const std::string& frame) {
    append(o, raw(seq));
    const uint64_t t = to_time(stamp);
    append(o, raw((uint32_t)(t & 0xffffffffu))); append(o, raw((uint32_t)(t >> 32)));
    append(o, raw((uint32_t)frame.size())); o.insert(o.end(), frame.begin(), frame.end());
  }
  static std::vector<uint8_t> encode_imu(uint32_t seq, double stamp, const double acc[3], const double gyr[3]) {
    std::vector<uint8_t> o;
    put_header(o, seq, stamp, "imu_link");
    const double q[4] = {0, 0, 0, 1}, z9[9] = {0};
    for (double v : q) append(o, raw(v));
    for (double v : z9) append(o, raw(v));
    for (int i = 0; i < 3; ++i) append(o, raw(gyr[i]));
    for (double v : z9) append(o, raw(v));
    for (int i = 0; i < 3; ++i) append(o, raw(acc[i]));
    for (double v : z9) append(o, raw(v));
    return o;
  }
  static std::vector<uint8_t> encode_cloud_xyzi(uint32_t seq, double stamp, const std::string& frame, const Cloud& cl) {
    std::vector<uint8_t> o;
    put_header(o, seq, stamp, frame);
    append(o, raw((uint32_t)1)); append(o, raw((uint32_t)cl.size()));
    append(o, raw((uint32_t)4));
    const char* names[4] = {"x", "y", "z", "intensity"};
    const uint32_t offs[4] = {0, 4, 8, 16};  // the PCL PointXYZI wire layout: 32-byte step
    for (int k = 0; k < 4; ++k) {
      const std::string nm = names[k];
      append(o, raw((uint32_t)nm.size())); o.insert(o.end(), nm.begin(), nm.end());
      append(o, raw(offs[k])); o.push_back(7); append(o, raw((uint32_t)1));
    }
    o.push_back(0);
    append(o, raw((uint32_t)32)); append(o, raw((uint32_t)(32 * cl.size())));
    append(o, raw((uint32_t)(32 * cl.size())));
    const uint8_t* p = reinterpret_cast<const uint8_t*>(cl.points.data());
    o.insert(o.end(), p, p + 32 * cl.size());
    o.push_back(1);
    return o;
  }
  static std::vector<uint8_t> encode_cloud_info(uint32_t seq, double stamp, const CloudInfo& ci) {
    std::vector<uint8_t> o;
    put_header(o, seq, stamp, "base_link");
    put_array(o, ci.startRingIndex); put_array(o, ci.endRingIndex);
    append(o, raw(ci.startOrientation)); append(o, raw(ci.endOrientation)); append(o, raw(ci.orientationDiff));
    put_array(o, ci.segmentedCloudGroundFlag); put_array(o, ci.segmentedCloudColInd); put_array(o, ci.segmentedCloudRange);
    return o;
  }

 private:
  struct Conn { uint32_t id; std::string topic, type, md5, def; };
  struct IndexEntry { uint64_t t; uint32_t off; };
  struct ChunkInfo { uint64_t pos, t0, t1; std::map<uint32_t, uint32_t> counts; };
  FILE* f_ = nullptr;
  std::vector<Conn> conns_;
  std::map<uint32_t, bool> announced_;
  std::vector<uint8_t> chunk_;
  std::map<uint32_t, std::vector<IndexEntry>> index_;
  std::vector<ChunkInfo> chunk_infos_;
  uint64_t t0_ = ~0ull, t1_ = 0;

  static uint64_t to_time(double t) {
    const uint32_t s = (uint32_t)t;
    uint32_t ns = (uint32_t)((t - (double)s) * 1e9 + 0.5);
    if (ns >= 1000000000u) ns = 999999999u;
    return (uint64_t)s | ((uint64_t)ns << 32);
  }
  template <typename T>
  static std::string raw(T v) { return std::string(reinterpret_cast<const char*>(&v), sizeof(T)); }
  static void append(std::vector<uint8_t>& o, const std::string& s) { o.insert(o.end(), s.begin(), s.end()); }
  template <typename T>
  static void put_array(std::vector<uint8_t>& o, const std::vector<T>& v) {
    append(o, raw((uint32_t)v.size()));
    const uint8_t* p = reinterpret_cast<const uint8_t*>(v.data());
    o.insert(o.end(), p, p + v.size() * sizeof(T));
  }
  static void put_field(std::vector<uint8_t>& h, const std::string& name, const std::string& value) {
    append(h, raw((uint32_t)(name.size() + 1 + value.size())));
    append(h, name); h.push_back('='); append(h, value);
  }
  static void put_record(std::vector<uint8_t>& o, const std::vector<uint8_t>& h, const std::vector<uint8_t>& d) {
    append(o, raw((uint32_t)h.size())); o.insert(o.end(), h.begin(), h.end());
    append(o, raw((uint32_t)d.size())); o.insert(o.end(), d.begin(), d.end());
  }
  static void append_connection(std::vector<uint8_t>& o, const Conn& c) {
    std::vector<uint8_t> h, d;
    put_field(h, "op", std::string(1, (char)0x07));
    put_field(h, "conn", raw(c.id));
    put_field(h, "topic", c.topic);
    put_field(d, "topic", c.topic);
    put_field(d, "type", c.type);
    put_field(d, "md5sum", c.md5);
    put_field(d, "message_definition", c.def);
    put_record(o, h, d);
  }
  void write_bag_header(uint64_t index_pos, uint32_t nconn, uint32_t nchunk) {
    std::vector<uint8_t> h;
    put_field(h, "op", std::string(1, (char)0x03));
    put_field(h, "index_pos", raw(index_pos));
    put_field(h, "conn_count", raw(nconn));
    put_field(h, "chunk_count", raw(nchunk));
    std::vector<uint8_t> d(4096 - 4 - h.size() - 4, (uint8_t)' ');
    std::vector<uint8_t> rec;
    put_record(rec, h, d);
    std::fwrite(rec.data(), 1, rec.size(), f_);
  }
  void flush_chunk() {
    if (chunk_.empty()) return;
    ChunkInfo ci;
    ci.pos = (uint64_t)std::ftell(f_); ci.t0 = t0_; ci.t1 = t1_;
    std::vector<uint8_t> h, rec;
    put_field(h, "op", std::string(1, (char)0x05));
    put_field(h, "compression", "none");
    put_field(h, "size", raw((uint32_t)chunk_.size()));
    put_record(rec, h, chunk_);
    for (const auto& kv : index_) {  // index data records follow their chunk
      std::vector<uint8_t> ih, id;
      put_field(ih, "op", std::string(1, (char)0x04));
      put_field(ih, "ver", raw((uint32_t)1));
      put_field(ih, "conn", raw(kv.first));
      put_field(ih, "count", raw((uint32_t)kv.second.size()));
      for (const auto& e : kv.second) { append(id, raw(e.t)); append(id, raw(e.off)); }
      put_record(rec, ih, id);
      ci.counts[kv.first] = (uint32_t)kv.second.size();
    }
    std::fwrite(rec.data(), 1, rec.size(), f_);
    chunk_infos_.push_back(ci);
    chunk_.clear(); index_.clear(); announced_.clear();
    t0_ = ~0ull; t1_ = 0;
  }
};

}  // namespace rosbag
}  // namespace lins

#endif  // LINS_HOST_ROSBAG_READER_HPP_
