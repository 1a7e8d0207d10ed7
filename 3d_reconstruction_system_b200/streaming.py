"""Host-side streaming for the sequence drivers: a frame batch is decoded on the native host
thread pool into PINNED memory while the batch before it is on the GPU, and finished text leaves for the file on a writer
thread while the next batch is formatted -- decode, upload / kernels / read-back and file I/O overlap, and host memory is
bounded by two batches whatever the length of the sequence.  (The reference keeps every point of the sequence in three
Python lists, transfer/camera_to_world.py:144-146.)"""
import ctypes as C
import queue
import threading

import numpy as np

from . import _lib, formats


class PinnedBuffer:
    """Grow-only page-locked host buffer (r3d_host_alloc) handed out as numpy views."""

    def __init__(self):
        self.lib = _lib.load()
        self.ptr, self.nbytes = None, 0

    def view(self, shape, dtype):
        dt = np.dtype(dtype)
        need = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        if need > self.nbytes:
            self.free()
            want = max(need + need // 8, 1 << 20)
            self.ptr = self.lib.r3d_host_alloc(want)
            if not self.ptr:
                # page-locking failed (no CUDA device in this process: the decode-only tools, or the limit on locked memory):
                # plain memory works everywhere a pinned buffer does, the copies are just slower
                self._plain = np.empty(want, dtype=np.uint8)
                self.nbytes = want
            else:
                self._plain = None
                self.nbytes = want
        if getattr(self, "_plain", None) is not None:
            return self._plain[:need].view(dt).reshape(shape)
        ct = (C.c_uint8 * max(need, 1)).from_address(self.ptr)
        return np.frombuffer(ct, dtype=dt, count=need // dt.itemsize).reshape(shape)

    def free(self):
        if self.ptr:
            self.lib.r3d_host_free(self.ptr)
        self.ptr, self.nbytes, self._plain = None, 0, None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchDecoder:
    """Iterates over (stack, n_frames) prefixes of `paths` like formats.read_frame_batch, decoding batch k + 1 on a worker
    thread (the native decoder releases the GIL) into one of two pinned stacks while the caller works on batch k.  The
    stack yielded for batch k is valid until the caller asks for the next batch (its buffer then goes back to the decoder)."""

    def __init__(self, paths, mode="gray", frames_per_batch=64):
        self.paths, self.mode, self.fpb = list(paths), mode, int(frames_per_batch)
        self.bufs = [PinnedBuffer(), PinnedBuffer()]
        self.free = threading.Semaphore(2)
        self.q = queue.Queue(maxsize=2)
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        k, turn = 0, 0
        try:
            while k < len(self.paths):
                self.free.acquire()
                buf = self.bufs[turn]
                stack, used = formats.read_frame_batch(self.paths[k:k + self.fpb], self.mode, max_frames=self.fpb, alloc=buf.view)
                self.q.put((stack, used))
                k += used
                turn ^= 1
            self.q.put(None)
        except BaseException as exc:          # hand the error to the consumer
            self.q.put(exc)

    def __iter__(self):
        first = True
        while True:
            if not first:
                self.free.release()            # the batch before the one being handed out now is finished with
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            first = False
            yield item

    def close(self):
        for b in self.bufs:
            b.free()


class AsyncFileWriter:
    """Appends byte blocks to an open binary file on a writer thread.  write(block, release) queues a block (a memoryview
    into a buffer the caller must not touch until `release` is called by the writer)."""

    def __init__(self, f):
        self.f = f
        self.q = queue.Queue(maxsize=4)
        self.err = None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            block, release = item
            try:
                if self.err is None:
                    self.f.write(block)
            except BaseException as exc:
                self.err = exc
            finally:
                if release is not None:
                    release()

    def write(self, block, release=None):
        if self.err is not None:
            raise self.err
        self.q.put((block, release))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.err is not None:
            raise self.err


class TextSlots:
    """Two pinned text buffers used alternately: the GPU formats rows into one while the writer thread drains the other."""

    def __init__(self):
        self.bufs = [PinnedBuffer(), PinnedBuffer()]
        self.sems = [threading.Semaphore(1), threading.Semaphore(1)]
        self.turn = 0

    def acquire(self, nbytes):
        i = self.turn
        self.turn ^= 1
        self.sems[i].acquire()
        return i, self.bufs[i].view((int(nbytes),), np.uint8)

    def releaser(self, i):
        return self.sems[i].release

    def regrow(self, i, nbytes):
        return self.bufs[i].view((int(nbytes),), np.uint8)

    def close(self):
        for s in self.sems:           # wait for the writer to be done with both
            s.acquire()
            s.release()
        for b in self.bufs:
            b.free()
