"""``--profile t``: host-side profile of a training run, what the reference wraps around ``run_full_training``
(Sol/Utilities/Profiler.py:1-16 used at Sol/Model/simulation_controller.py:111-117): cProfile over the block, statistics
sorted by cumulative time on exit.  On the GPU path the host profile mostly shows where the Python driver waits for the
device; kernel-level numbers come from ``bench.py`` / ncu (profiles/)."""
import cProfile
import io
import pstats


class Profiler:
    def __init__(self, print_fn=print, top: int = 60):
        self.profile, self.print_fn, self.top = cProfile.Profile(), print_fn, top
        self.report = ""

    def __enter__(self):
        self.profile.enable()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.profile.disable()
        buf = io.StringIO()
        st = pstats.Stats(self.profile, stream=buf).sort_stats("cumtime")
        st.print_stats(self.top)
        # the same ranking restricted to this package's own functions (in short runs torch's lazy imports fill the list above)
        buf.write("---- functions of drl_dronenavigation_b200 only ----\n")
        st.print_stats(r"drl[-_]dronenavigation", self.top)
        self.report = buf.getvalue()
        self.print_fn(self.report)
        return False
