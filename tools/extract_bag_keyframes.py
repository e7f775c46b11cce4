"""Extracts keyframes (odometry pose + laser scan) of one robot from a reference bag file, without
ROS: a minimal ROSBAG V2.0 reader (uncompressed chunks).

    python tools/extract_bag_keyframes.py /tmp/bags/2robots-hospital.bag robot_0 80 tests/golden/bag_2robots_robot0_kf80.npz

Keyframe rule of the reference's main loop (src/srslam.cpp:195-201): a new keyframe when the
odometry moved more than 0.25 m or turned more than pi/4 since the last one. The reference samples
the latest odometry / scan at 10 Hz of wall-clock time; offline the rule is evaluated at every
odometry message and paired with the most recent scan (appendix A: a replay must fix a sampling
rule; parity is CPU oracle vs GPU on the identical replay). The fixture holds float64 odometry
poses (x, y, yaw), float32 ranges as stored in the bag, the scan geometry, and per keyframe the bag
time (s) and the robot's latest ground-truth pose (the multi-robot replay needs both: message
cadence and the simulated communication range, graph_comm.cpp:66-68,152).
The bags are under /root/reference/bagfiles/*.tar.gz (extract first). Test-data tooling only."""
import math
import struct
import sys

import numpy as np


def records(buf, pos, end):
    while pos < end:
        hlen = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        hdr = {}
        hend = pos + hlen
        while pos < hend:
            flen = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
            field = buf[pos:pos + flen]
            pos += flen
            k, _, v = field.partition(b"=")
            hdr[k.decode()] = v
        dlen = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        yield hdr, pos, dlen
        pos += dlen


def messages(path):
    """(topic, time_ns, payload) of every message, in file order."""
    buf = open(path, "rb").read()
    assert buf.startswith(b"#ROSBAG V2.0\n")
    topics = {}
    for hdr, pos, dlen in records(buf, 13, len(buf)):
        op = hdr["op"][0]
        if op == 0x05:  # chunk
            assert hdr["compression"] == b"none"
            for h2, p2, d2 in records(buf, pos, pos + dlen):
                op2 = h2["op"][0]
                if op2 == 0x07:
                    topics[struct.unpack("<I", h2["conn"])[0]] = h2["topic"].decode()
                elif op2 == 0x02:
                    sec, nsec = struct.unpack("<II", h2["time"])
                    yield topics[struct.unpack("<I", h2["conn"])[0]], sec * 10**9 + nsec, buf[p2:p2 + d2]
        elif op == 0x07:
            topics[struct.unpack("<I", hdr["conn"])[0]] = hdr["topic"].decode()


def skip_header(b, pos):  # std_msgs/Header: seq, stamp, frame_id
    pos += 12
    n = struct.unpack_from("<I", b, pos)[0]
    return pos + 4 + n


def parse_scan(b):
    pos = skip_header(b, 0)
    amin, amax, ainc, tinc, stime, rmin, rmax = struct.unpack_from("<7f", b, pos)
    pos += 28
    n = struct.unpack_from("<I", b, pos)[0]
    pos += 4
    ranges = np.frombuffer(b, dtype="<f4", count=n, offset=pos).copy()
    return amin, ainc, rmax, ranges


def parse_odom(b):
    pos = skip_header(b, 0)
    n = struct.unpack_from("<I", b, pos)[0]  # child_frame_id
    pos += 4 + n
    x, y, z, qx, qy, qz, qw = struct.unpack_from("<7d", b, pos)
    yaw = math.atan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz))
    return x, y, yaw


def main():
    path, robot, count, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    odom_topic, scan_topic = "/%s/odom" % robot, "/%s/base_scan" % robot
    gt_topic = "/%s/base_pose_ground_truth" % robot
    last_scan = None
    geom = None
    key_odom, key_ranges, key_time, key_gt = [], [], [], []
    last = None
    last_gt = (float("nan"),) * 3
    for topic, t, payload in messages(path):
        if topic == gt_topic:
            last_gt = parse_odom(payload)     # nav_msgs/Odometry as well
        elif topic == scan_topic:
            amin, ainc, rmax, ranges = parse_scan(payload)
            geom = (amin, ainc, rmax)
            last_scan = ranges
        elif topic == odom_topic and last_scan is not None:
            x, y, yaw = parse_odom(payload)
            if last is None:
                take = True
            else:
                dx, dy = x - last[0], y - last[1]
                dth = (yaw - last[2] + math.pi) % (2 * math.pi) - math.pi
                take = math.hypot(dx, dy) > 0.25 or abs(dth) > math.pi / 4
            if take:
                last = (x, y, yaw)
                key_odom.append(last)
                key_ranges.append(last_scan)
                key_time.append(t * 1e-9)
                key_gt.append(last_gt)
                if len(key_odom) >= count:
                    break
    np.savez_compressed(out, odom=np.array(key_odom), ranges=np.array(key_ranges, dtype=np.float32),
                        time=np.array(key_time), gt=np.array(key_gt),
                        first_angle=np.float64(geom[0]), angular_step=np.float64(geom[1]),
                        max_range=np.float64(geom[2]), robot=robot, source=path.split("/")[-1])
    print("keyframes", len(key_odom), "beams", key_ranges[0].shape, "geom", geom, "->", out)


if __name__ == "__main__":
    main()
