"""Ray-segment logger (reference: debug/ray_logger.py:1-16, used by main.py:66-85).

Same container and the same two methods as the reference's RayLogger -- ``points`` (list of
3-vectors), ``lines`` (index pairs into ``points``), ``colors`` (one rgb per line) -- so code that
hands it to an Open3D ``LineSet`` (main.py:80-83) keeps working.  Here the segments come from the
device: ``core.tracing.path_tracing(ray, scene, ray_logger)`` runs the wavefront integrator with
the C ABI's path-segment log switched on (``prt_set_path_log``) and appends every traced segment.
``write_obj`` replaces the Open3D window with a Wavefront OBJ of line elements.
"""
import numpy as np

# line colour per segment kind: path segments by bounce (camera ray first), light connections yellow
BOUNCE_COLORS = [[1.0, 0.0, 0.0], [0.0, 0.6, 0.0], [0.0, 0.3, 1.0], [0.6, 0.0, 0.8], [0.3, 0.3, 0.3]]
LIGHT_COLOR = [1.0, 0.8, 0.0]


class RayLogger:
    def __init__(self):
        self.points = []
        self.lines = []
        self.colors = []
        self.kinds = []   # bounce index, or -1 for a light connection (not in the reference)
        self.paths = []   # path id of each line (ray index * samples + sample)

    def add(self, ray, t=5, color=[1, 0, 0]):
        self.add_line(ray.position, ray.position + t * ray.direction, color)

    def add_line(self, p1, p2, color=[1, 0, 0], kind=0, path=0):
        self.points.extend([np.asarray(p1, np.float64), np.asarray(p2, np.float64)])
        self.lines.append([len(self.points) - 2, len(self.points) - 1])
        self.colors.append(list(color))
        self.kinds.append(int(kind))
        self.paths.append(int(path))

    def add_device_segments(self, records):
        """records: float32 [n, 8] prt_segment rows (p0, bits(kind), p1, bits(path)), any order.
        Appended sorted by (path, light connection last, bounce) so the log is deterministic."""
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 8)
        kind = rec[:, 3].copy().view(np.int32)
        path = rec[:, 7].copy().view(np.uint32)
        order = np.lexsort((kind, kind < 0, path))
        for i in order:
            k = int(kind[i])
            color = LIGHT_COLOR if k < 0 else BOUNCE_COLORS[min(k, len(BOUNCE_COLORS) - 1)]
            self.add_line(rec[i, 0:3], rec[i, 4:7], color, k, int(path[i]))

    def write_obj(self, filename):
        with open(filename, "w") as f:
            for p in self.points:
                f.write(f"v {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}\n")
            for a, b in self.lines:
                f.write(f"l {a + 1} {b + 1}\n")
