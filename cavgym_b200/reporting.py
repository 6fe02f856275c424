"""Episode / run accounting with the reference's definitions (reporting.py:120-269): an episode is "interesting"
when a tester won (info['winner'] > 0), its score is minus the summed liveness of the testers, and a run reports the
95 % confidence intervals of timesteps, runtime and score over the interesting episodes.  The batched engine keeps
the sufficient statistics (counts, sums, sums of squares) on the device (cavgym_stats); `RunSummary.from_stats`
turns them into the same figures."""
import logging
import math
import sys
from dataclasses import dataclass
from enum import Enum


class Verbosity(Enum):
    INFO = "info"
    DEBUG = "debug"
    SILENT = "silent"

    def __str__(self):
        return self.value


def pretty_float(value, decimal_places=2):
    return f"{round(value, decimal_places):g}"


def pretty_str_list(items):
    return f"[{', '.join(map(str, items))}]"


def get_console(verbosity):
    console = logging.getLogger("cavgym_b200.console")
    console.setLevel({Verbosity.DEBUG: logging.DEBUG, Verbosity.INFO: logging.INFO, Verbosity.SILENT: logging.WARNING}[verbosity])
    if not console.handlers:
        formatter = logging.Formatter("%(levelname)-7s %(relativeCreated)-7d %(message)s")
        for stream, accept in ((sys.stdout, lambda r: r.levelno <= logging.INFO), (sys.stderr, lambda r: r.levelno > logging.INFO)):
            handler = logging.StreamHandler(stream)
            handler.addFilter(accept)
            handler.setFormatter(formatter)
            console.addHandler(handler)
    return console


def get_file_logger(path):
    import pathlib
    logger = logging.getLogger(path)
    if not logger.handlers:
        pathlib.Path(path).parent.mkdir(parents=True, exist_ok=True)
        handler = logging.FileHandler(path)
        handler.setFormatter(logging.Formatter("%(message)s"))
        logger.addHandler(handler)
        logger.setLevel(logging.INFO)
    return logger


get_episode_file_logger = get_run_file_logger = get_agent_file_logger = get_file_logger


@dataclass(frozen=True)
class Interval:
    value: float
    error: float

    def __iter__(self):
        yield self.value
        yield self.error

    def pretty(self, **kwargs):
        return f"{pretty_float(self.value, **kwargs)} ± {pretty_float(self.error, **kwargs)}"


def _t_interval(count, mean, sem, alpha=0.05):
    from scipy import stats
    return Interval(value=mean, error=sem * stats.t.isf(alpha / 2, count - 1))


def confidence_interval(data, alpha=0.05):
    """Mean and half-width of the (1 - alpha) Student-t interval; NaN with fewer than two points."""
    n = len(data)
    if n <= 1:
        return Interval(float("nan"), float("nan"))
    mean = sum(data) / n
    variance = sum((x - mean) ** 2 for x in data) / (n - 1)
    return _t_interval(n, mean, math.sqrt(variance / n), alpha)


def interval_from_sums(count, total, total_sq, alpha=0.05):
    """Same interval from the sufficient statistics the device accumulates (count, sum x, sum x^2)."""
    if count <= 1:
        return Interval(float("nan"), float("nan"))
    mean = total / count
    variance = max(0.0, (total_sq - count * mean * mean) / (count - 1))
    return _t_interval(count, mean, math.sqrt(variance / count), alpha)


@dataclass(frozen=True)
class TimeResults:
    timesteps: int
    start_time: float
    end_time: float
    resolution: float

    def runtime(self):
        return (self.end_time - self.start_time) * 1000

    def simulation_speed(self):
        return (self.timesteps * self.resolution * 1000) / self.runtime()


@dataclass(frozen=True)
class EpisodeResults:
    index: int
    time: TimeResults
    completed: bool
    interesting: bool
    score: float

    def console_message(self):
        status = "completed" if self.completed else "terminated"
        test = f"interesting test with score {self.score}" if self.interesting else "uninteresting test"
        return (f"episode {self.index} {status} after {self.time.timesteps} timestep(s) in "
                f"{pretty_float(self.time.runtime(), decimal_places=0)} ms (*{pretty_float(self.time.simulation_speed())} real-time), {test}")

    def file_message(self):
        return f"{self.index},{self.time.timesteps},{self.time.runtime()},{1 if self.interesting else 0},{self.score}"


def analyse_episode(index, start_time, end_time, timesteps, env_info, run_config, env):
    assert 1 <= timesteps <= run_config.max_timesteps
    interesting = 'winner' in env_info and env_info['winner'] > 0
    score = -sum(env.episode_liveness[1:]) if interesting else float("nan")
    return EpisodeResults(index, TimeResults(timesteps, start_time, end_time, env.time_resolution),
                          completed=timesteps == run_config.max_timesteps, interesting=interesting, score=score)


@dataclass(frozen=True)
class RunSummary:
    """analyse_run (reporting.py:246-269) and RunResults' two messages (reporting.py:219-224).  The three intervals are over
    the INTERESTING episodes only, as in the reference; the per-episode runtime has no meaning for a batch of concurrent
    environments, so that interval is NaN there (its two run_log columns are kept: the line has the reference's 10 fields)."""
    episodes: int
    timesteps: int
    runtime_ms: float
    resolution: float
    interesting: int
    confidence_timesteps: Interval
    confidence_score: Interval
    confidence_runtime: Interval = Interval(float("nan"), float("nan"))

    @classmethod
    def from_episodes(cls, episode_data, start_time, end_time, resolution):
        hits = [row for row in episode_data if row.interesting]
        return cls(len(episode_data), sum(row.time.timesteps for row in episode_data), (end_time - start_time) * 1000, resolution,
                   len(hits), confidence_interval([row.time.timesteps for row in hits]), confidence_interval([row.score for row in hits]),
                   confidence_interval([row.time.runtime() for row in hits]))

    @classmethod
    def from_stats(cls, stats, runtime_ms, resolution):
        """`stats` is the dict of cavgym_stats (summed over GPUs): the device keeps count, sum and sum of squares of the
        timesteps and of the score over the interesting episodes (score_episode, transition.cuh)."""
        return cls(stats["episodes"], stats["env_steps"], runtime_ms, resolution, stats["interesting"],
                   interval_from_sums(stats["interesting"], stats["sum_t_interesting"], stats["sum_t2_interesting"]),
                   interval_from_sums(stats["interesting"], stats["sum_score"], stats["sum_score2"]))

    def simulation_speed(self):
        return (self.timesteps * self.resolution * 1000) / self.runtime_ms

    def console_message(self):
        tests = (f"{self.interesting} interesting test(s) with {self.confidence_timesteps.pretty(decimal_places=0)} timestep(s), "
                 f"{self.confidence_runtime.pretty(decimal_places=0)} ms runtime, and "
                 f"{self.confidence_score.pretty(decimal_places=0)} score") if self.interesting > 0 else "no interesting test(s)"
        return (f"run completed after {self.episodes} episode(s) and {self.timesteps} timestep(s) in "
                f"{pretty_float(self.runtime_ms, decimal_places=0)} ms (*{pretty_float(self.simulation_speed())} real-time), {tests}")

    def file_message(self):
        return (f"{self.episodes},{self.timesteps},{self.runtime_ms},{self.interesting},{self.confidence_timesteps.value},"
                f"{self.confidence_timesteps.error},{self.confidence_runtime.value},{self.confidence_runtime.error},"
                f"{self.confidence_score.value},{self.confidence_score.error}")
