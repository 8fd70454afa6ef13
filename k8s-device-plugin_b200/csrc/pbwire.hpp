// pbwire.hpp -- minimal protobuf (proto3) wire encoder for the v1beta1 device-plugin
// messages the hot path emits (vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto).
// Field order ascending and zero values omitted, i.e. byte-identical to the reference's
// gogo-protobuf Marshal for these messages.
#pragma once
#include <cstdint>
#include <string>

namespace b2dp {
namespace pb {

inline void varint(std::string& o, uint64_t v) {
    while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
    o.push_back((char)v);
}
inline void tag(std::string& o, int field, int wire) { varint(o, ((uint64_t)field << 3) | (uint64_t)wire); }
inline void bytes_field(std::string& o, int field, const std::string& b) {
    tag(o, field, 2); varint(o, b.size()); o.append(b);
}
// proto3 string: omitted when empty
inline void string_field(std::string& o, int field, const std::string& s) { if (!s.empty()) bytes_field(o, field, s); }
// proto3 int64: omitted when zero; negative values take 10 bytes
inline void int64_field(std::string& o, int field, int64_t v) { if (v) { tag(o, field, 0); varint(o, (uint64_t)v); } }

// Device{ID=1, health=2, topology=3{nodes=1[{ID=1}]}}, api.proto:96-118
inline void encode_device(std::string& o, const std::string& id, const std::string& health, int64_t numa) {
    std::string numa_node; int64_field(numa_node, 1, numa);
    std::string topo; bytes_field(topo, 1, numa_node);
    std::string dev;
    string_field(dev, 1, id); string_field(dev, 2, health); bytes_field(dev, 3, topo);
    bytes_field(o, 1, dev);  // ListAndWatchResponse.devices = 1
}
// DeviceSpec{container_path=1, host_path=2, permissions=3}, api.proto:199-210
inline void encode_devspec(std::string& o, int field, const std::string& container, const std::string& host,
                           const std::string& perms) {
    std::string d;
    string_field(d, 1, container); string_field(d, 2, host); string_field(d, 3, perms);
    bytes_field(o, field, d);
}

}  // namespace pb
}  // namespace b2dp
