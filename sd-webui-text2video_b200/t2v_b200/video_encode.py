"""Minimal stand-in for the reference's PNG + ffmpeg packaging (process_modelscope.py:224-254: frames -> vid.mp4 ->
"data:video/mp4;base64,..."), so that `process_modelscope` returns `list[str]` data URLs like the reference when the webui's
own `ffmpeg_stitch_video` is not plugged in.

  * if an `ffmpeg` binary is on PATH the clip is encoded to H.264 mp4 through a pipe (same container / URL prefix as the
    reference, no temporary PNG files);
  * otherwise the frames are wrapped, uncompressed, in a RIFF AVI ('DIB ' 24-bit BGR, bottom-up rows) written in pure Python
    and returned as "data:video/avi;base64,..." -- every browser-side consumer that only checks for a data URL keeps working,
    and the bytes are exact.
Host-side packaging only: nothing here touches the GPU path."""
import base64
import shutil
import struct
import subprocess

import numpy as np


def _avi_bytes(frames, fps):
    h, w = frames[0].shape[:2]
    row = (w * 3 + 3) & ~3
    img_size = row * h
    n = len(frames)

    def chunk(fourcc, data):
        return fourcc + struct.pack('<I', len(data)) + data + (b'\x00' if len(data) & 1 else b'')

    def lst(kind, data):
        return b'LIST' + struct.pack('<I', len(data) + 4) + kind + data
    avih = struct.pack('<IIIIIIIIIIIIII', int(1e6 / fps), img_size * int(fps), 0, 0x10, n, 0, 1, img_size, w, h, 0, 0, 0, 0)
    strh = struct.pack('<4s4sIHHIIIIIIIIhhhh', b'vids', b'DIB ', 0, 0, 0, 0, 1, int(fps), 0, n, img_size, 0xFFFFFFFF, 0, 0, 0, w, h)
    strf = struct.pack('<IiiHHIIiiII', 40, w, h, 1, 24, 0, img_size, 0, 0, 0, 0)
    hdrl = lst(b'hdrl', chunk(b'avih', avih) + lst(b'strl', chunk(b'strh', strh) + chunk(b'strf', strf)))
    movi_data, index, off = b'', b'', 4
    pad = np.zeros((h, row - w * 3), dtype=np.uint8)
    for f in frames:
        rows = np.ascontiguousarray(f[::-1].reshape(h, w * 3))              # bottom-up, BGR as delivered by infer()
        data = np.concatenate([rows, pad], axis=1).tobytes() if pad.shape[1] else rows.tobytes()
        c = chunk(b'00db', data)
        index += struct.pack('<4sIII', b'00db', 0x10, off, len(data))
        off += len(c)
        movi_data += c
    body = hdrl + lst(b'movi', movi_data) + chunk(b'idx1', index)
    return b'RIFF' + struct.pack('<I', len(body) + 4) + b'AVI ' + body


def default_video_encoder(frames, args=None):
    """list of HxWx3 uint8 BGR frames -> data-URL string (mp4 through ffmpeg when available, else uncompressed AVI)."""
    fps = float(getattr(args, 'fps', 15) or 15)
    frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
    h, w = frames[0].shape[:2]
    ff = shutil.which('ffmpeg')
    if ff is not None and h % 2 == 0 and w % 2 == 0:
        cmd = [ff, '-loglevel', 'error', '-f', 'rawvideo', '-pix_fmt', 'bgr24', '-s', f'{w}x{h}', '-r', str(fps), '-i', '-', '-an',
               '-c:v', 'libx264', '-pix_fmt', 'yuv420p', '-crf', str(getattr(args, 'ffmpeg_crf', 17)), '-preset',
               str(getattr(args, 'ffmpeg_preset', 'slow')), '-movflags', 'frag_keyframe+empty_moov', '-f', 'mp4', '-']
        try:
            r = subprocess.run(cmd, input=b''.join(f.tobytes() for f in frames), capture_output=True, timeout=600)
            if r.returncode == 0 and r.stdout:
                return 'data:video/mp4;base64,' + base64.b64encode(r.stdout).decode()
        except Exception:
            pass
    return 'data:video/avi;base64,' + base64.b64encode(_avi_bytes(frames, fps)).decode()
