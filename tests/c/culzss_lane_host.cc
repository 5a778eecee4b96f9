// Host build of the lane-serial CULZSS encoders (csrc/culzss_lane.cuh) for the CPU test
// tests/test_culzss_lane_cpu.py: the same code a GPU lane runs, STRIDE = 1.
#include "culzss_lane.cuh"

using namespace b200lc::lzss_lane;

struct HostIO {
    const u8 *src;
    u8 *dst;
    bool any(bool b) const { return b; }
    void load(u32 off, Chunk32 &c) const { memcpy(c.w, src + off, 32); }
    u32 bytes4(u32 off) const
    {
        u32 v = 0;
        for (u32 j = 0; j < 4; ++j)
            if (off + j < kPacket) v |= (u32)src[off + j] << (8 * j);
        return v;
    }
    void store(u32 off, const u32 (&x)[4]) const { memcpy(dst + off, x, 16); }
};

// Encodes npackets packets of 4096 bytes; packet k's bytes go to out + k * kSlotBytes.
// parity != 0: the reference's matches (bit-exact mode); 0: fast mode.
extern "C" void lane_encode_packets(const u8 *in, u32 npackets, u8 *out, u16 *sizes, u8 *last_group, int parity)
{
    for (u32 k = 0; k < npackets; ++k) {
        u32 column[kColumnWordsFast];
        HostIO io{in + (size_t)k * kPacket, out + (size_t)k * kSlotBytes};
        u32 lg = 0;
        sizes[k] = (u16)(parity ? encode_packet<1, true>(column, true, io, lg)
                                : encode_packet<1, false>(column, true, io, lg));
        last_group[k] = (u8)lg;
    }
}
