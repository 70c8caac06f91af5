"""Summarise the in-kernel time stamps of the fused reverse kernels (library built with `make TRACE=1`,
$ADEPT_CUDA_TRACE_FILE): medians of the phases, in microseconds (globaltimer ticks are nanoseconds).

  per-step kernel (8 stamps per launch of stream 0):  0 first CTA starts, 1/2 first/last CTA past griddepcontrol.wait,
      3 last CTA has its activity bytes, 4 last CTA finished its tasks, 6 last CTA done
  resident kernel (4 stamps per item of CTA 0):  0 wait begins, 1 wait ends, 2 activity bytes in, 3 item done
"""
import sys
import numpy as np

rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip()]
a = np.array(rows, dtype=np.float64)
med = lambda x: float(np.median(x)) / 1e3
if a.shape[1] == 4:
    a = a[a[:, 3] > 0]
    print("resident kernel, CTA 0, %d items" % len(a))
    print("  wait for the step above   %.2f us" % med(a[:, 1] - a[:, 0]))
    print("  records + activity bytes  %.2f us" % med(a[:, 2] - a[:, 1]))
    print("  tasks (row loads, stores) %.2f us" % med(a[:, 3] - a[:, 2]))
    print("  item to item              %.2f us" % med(a[1:, 0] - a[:-1, 0]))
    print("  whole trace               %.2f ms for %d items" % ((a[-1, 3] - a[0, 0]) / 1e6, len(a)))
else:
    a = a[(a[:, 6] > 0) & (a[:, 0] < 2 ** 63)]
    print("per-step kernel, stream 0, %d launches" % len(a))
    print("  first CTA start -> first past griddepcontrol.wait  %.2f us" % med(a[:, 1] - a[:, 0]))
    print("  first -> last CTA past the wait                     %.2f us" % med(a[:, 2] - a[:, 1]))
    print("  wait -> activity bytes in (last CTA)                %.2f us" % med(a[:, 3] - a[:, 2]))
    print("  activity bytes -> tasks done (last CTA)             %.2f us" % med(a[:, 4] - a[:, 3]))
    print("  tasks done -> kernel end (tails included)           %.2f us" % med(a[:, 6] - a[:, 4]))
    print("  launch to launch (wait to wait)                     %.2f us" % med(a[1:, 1] - a[:-1, 1]))
