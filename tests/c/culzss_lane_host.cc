// Host build of the lane-serial CULZSS fast-mode encoder (csrc/culzss_lane.cuh) for the CPU test
// tests/test_culzss_lane_cpu.py: the same code a GPU lane runs, STRIDE = 1.
#include "culzss_lane.cuh"

using namespace b200lc::lzss_lane;

// Encodes npackets packets of 4096 bytes; packet k's bytes go to out + k * kSlotBytes.
extern "C" void lane_encode_packets(const u8 *in, u32 npackets, u8 *out, u16 *sizes, u8 *last_group)
{
    for (u32 k = 0; k < npackets; ++k) {
        u32 column[kColumnWords];
        alignas(16) u8 pkt[kPacket];
        memcpy(pkt, in + (size_t)k * kPacket, kPacket);
        u8 *dst = out + (size_t)k * kSlotBytes;
        Lane<1> ln;
        ln.init(column, pkt, dst);
        for (;;) {
            while (ln.wants_input()) {
                u32 x[4];
                memcpy(x, ln.src + ln.hi, 16);
                ln.put_input(x[0], x[1], x[2], x[3]);
            }
            if (ln.p >= kPacket) break;
            ln.step();
            if (ln.has_output()) {
                u32 x[4];
                const u32 at = ln.flushed;
                ln.take_output(x[0], x[1], x[2], x[3]);
                memcpy(dst + at, x, 16);
            }
        }
        ln.finish();
        while (ln.flushed < ln.o) {
            u32 x[4];
            const u32 at = ln.flushed;
            ln.take_output(x[0], x[1], x[2], x[3]);
            memcpy(dst + at, x, 16);
        }
        sizes[k] = (u16)ln.o;
        last_group[k] = (u8)ln.last_group_bytes();
    }
}
