"""Dependency-free pcap framing in the layout the reference reads and writes.

vtkPacketFileWriter (vtkPacketFileWriter.cxx:41-54, 118-161) fabricates a 42-byte
Ethernet/IPv4/UDP header in front of each 1206-byte payload and dumps it through libpcap
(DLT_EN10MB): 24-byte global header, then 16 + 42 + 1206 = 1264-byte records
(vtkPacketFileReader.h:57-66).  vtkPacketFileReader::nextPacket strips the 42 bytes and turns
the record timeval into a ptime with timevalToPtime, which adds 8 hours
(type_defs.cxx:69-72); the same rule is applied here and by the GPU-side k_pcap_times.
"""
from __future__ import annotations

import struct

import numpy as np

GLOBAL_HEADER_BYTES = 24
RECORD_HEADER_BYTES = 16
NET_HEADER_BYTES = 42
PAYLOAD_BYTES = 1206
RECORD_BYTES = RECORD_HEADER_BYTES + NET_HEADER_BYTES + PAYLOAD_BYTES   # 1264
PAYLOAD_OFFSET = GLOBAL_HEADER_BYTES + RECORD_HEADER_BYTES + NET_HEADER_BYTES  # 82
TZ_SHIFT_US = 8 * 3600 * 1_000_000


def lidar_net_header():
    """Broadcast Ethernet + IPv4 (192.168.1.200 -> 255.255.255.255) + UDP 2368 -> 2368."""
    eth = bytes([0xff] * 6) + bytes([0x60, 0x76, 0x88, 0x00, 0x00, 0x00]) + bytes([0x08, 0x00])
    ip = bytes([0x45, 0x00, 0x04, 0xd2, 0x00, 0x00, 0x40, 0x00, 0xff, 0x11, 0xb4, 0xaa,
                0xc0, 0xa8, 0x01, 0xc8, 0xff, 0xff, 0xff, 0xff])
    udp = struct.pack(">HHHH", 2368, 2368, 8 + PAYLOAD_BYTES, 0)
    h = eth + ip + udp
    assert len(h) == NET_HEADER_BYTES
    return np.frombuffer(h, dtype=np.uint8)


def write_pcap_image(payloads, t_us):
    """File image (uint8 array) for n payloads of 1206 bytes with reader-side times t_us."""
    payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
    n = payloads.shape[0]
    assert payloads.shape[1] == PAYLOAD_BYTES
    t = np.asarray(t_us, dtype=np.int64) - TZ_SHIFT_US
    img = np.zeros(GLOBAL_HEADER_BYTES + n * RECORD_BYTES, dtype=np.uint8)
    img[:GLOBAL_HEADER_BYTES] = np.frombuffer(
        struct.pack("<IHHiIII", 0xa1b2c3d4, 2, 4, 0, 0, 65535, 1), dtype=np.uint8)
    rec = img[GLOBAL_HEADER_BYTES:].reshape(n, RECORD_BYTES)
    hdr = np.zeros((n, 4), dtype="<u4")
    hdr[:, 0] = (t // 1_000_000).astype(np.uint32)
    hdr[:, 1] = (t % 1_000_000).astype(np.uint32)
    hdr[:, 2] = NET_HEADER_BYTES + PAYLOAD_BYTES
    hdr[:, 3] = NET_HEADER_BYTES + PAYLOAD_BYTES
    rec[:, :RECORD_HEADER_BYTES] = hdr.view(np.uint8).reshape(n, 16)
    rec[:, RECORD_HEADER_BYTES:RECORD_HEADER_BYTES + NET_HEADER_BYTES] = lidar_net_header()[None, :]
    rec[:, RECORD_HEADER_BYTES + NET_HEADER_BYTES:] = payloads
    return img


def write_pcap(path, payloads, t_us):
    write_pcap_image(payloads, t_us).tofile(path)


def payload_view(img):
    """(view starting at the first payload, number of 1264-byte records) of a uniform image."""
    n = (img.shape[0] - GLOBAL_HEADER_BYTES) // RECORD_BYTES
    return img[PAYLOAD_OFFSET:], n


def read_pcap_image(img):
    """Generic reader: (payloads[n, 1206], t_us[n]) of the UDP records whose payload is 1206 B
    (HDLParser.cxx:982 drops every other size)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    magic = struct.unpack_from("<I", img, 0)[0]
    if magic != 0xa1b2c3d4:
        raise ValueError("not a little-endian microsecond pcap file")
    off = GLOBAL_HEADER_BYTES
    out, ts = [], []
    n = img.shape[0]
    while off + RECORD_HEADER_BYTES <= n:
        sec, usec, caplen, length = struct.unpack_from("<IIII", img, off)
        body = off + RECORD_HEADER_BYTES
        if body + caplen > n:
            break
        if length - NET_HEADER_BYTES == PAYLOAD_BYTES and caplen >= length:
            out.append(img[body + NET_HEADER_BYTES:body + NET_HEADER_BYTES + PAYLOAD_BYTES])
            ts.append((sec * 1_000_000 + usec) + TZ_SHIFT_US)
        off = body + caplen
    if not out:
        return np.zeros((0, PAYLOAD_BYTES), np.uint8), np.zeros(0, np.int64)
    return np.stack(out), np.asarray(ts, dtype=np.int64)


def read_pcap(path):
    return read_pcap_image(np.fromfile(path, dtype=np.uint8))
