"""Host-side streaming of the sequence drivers (no GPU): the prefetching batch decoder hands out the frames of a file list
in order, in equally shaped batches, each stack valid until the next batch is asked for; the writer thread appends
blocks in order and hands buffers back."""
import importlib
import threading

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
streaming = importlib.import_module("3d_reconstruction_system_b200.streaming")
formats = importlib.import_module("3d_reconstruction_system_b200.formats")


def test_batch_decoder_order_shapes_and_lifetime(tmp_path):
    rng = np.random.default_rng(3)
    paths, want = [], []
    shapes = [(20, 30)] * 7 + [(12, 40)] * 3 + [(20, 30)] * 5
    for k, (h, w) in enumerate(shapes):
        img = rng.integers(0, 65536, size=(h, w)).astype(np.uint16)
        p = tmp_path / ("%03d.png" % k)
        cv2.imwrite(str(p), img)
        paths.append(str(p))
        want.append(img)
    dec = streaming.BatchDecoder(paths, "raw", frames_per_batch=4)
    import time
    got, sizes = [], []
    for stack, used in dec:
        assert stack.shape[0] == used and stack.flags.c_contiguous
        snap = stack.copy()
        time.sleep(0.05)                                        # the decoder is busy with the NEXT batch meanwhile ...
        assert np.array_equal(stack, snap)                      # ... and must not touch the one handed out
        sizes.append(used)
        got.extend(np.array(f) for f in stack)
    dec.close()
    assert sizes == [4, 3, 3, 4, 1]                            # a change of shape ends a batch
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    # gray mode: uint8 (16-bit >> 8), like cv.imread(..., IMREAD_GRAYSCALE)
    dec = streaming.BatchDecoder(paths[:7], "gray", frames_per_batch=16)
    (stack, used), = list(dec)
    dec.close()
    assert used == 7 and stack.dtype == np.uint8 and np.array_equal(stack[3], (want[3] >> 8).astype(np.uint8))
    # a missing file surfaces in the consumer
    dec = streaming.BatchDecoder(paths[:2] + [str(tmp_path / "missing.png")], "raw", frames_per_batch=2)
    with pytest.raises(Exception):
        list(dec)
    assert formats.frame_pixels(paths[0]) == 600 and formats.frame_pixels(paths[8]) == 480


def test_async_writer_and_text_slots(tmp_path):
    slots = streaming.TextSlots()
    path = tmp_path / "out.bin"
    with open(path, "wb") as f:
        w = streaming.AsyncFileWriter(f)
        expect = b""
        for k in range(9):
            i, buf = slots.acquire(1000 + 37 * k)
            buf[:] = k
            n = 500 + 11 * k
            expect += bytes([k]) * n
            w.write(memoryview(buf)[:n], slots.releaser(i))
        w.close()
        slots.close()
    assert path.read_bytes() == expect
    assert threading.active_count() < 20
