"""Drop-in `analyser` module: the consumers of what `Solver.save()` + `Solver.solve()` leave on disk
(`<name>.npy` = float64 snapshot stack, `<name>.txt` = JSON run constants).

Same public surface and behaviour as the reference's analyser.py (FileLoader: analyser.py:7-84, DataCollector:
:87-159) so that reference scripts run unchanged, but organised around three module-level helpers.  Host-side numpy
only: this is outside the accelerated path; matplotlib is imported lazily because it is optional
on a GPU box.
"""
import json

import numpy as np

# the scalar run constants both classes mirror as attributes (written by Solver.save_to_json)
RUN_KEYS = ("dt", "ds", "length_x", "length_y", "stability", "omega_max", "end_time", "step_frequency")
FFT_POINTS = 2 ** 14


def _mpl():
    import matplotlib.pyplot as pyplot
    import matplotlib.animation as animation
    return pyplot, animation


def _adopt(obj, constants):
    """Mirror the scalar run constants of a `.txt` file as attributes of `obj`."""
    for key in RUN_KEYS:
        setattr(obj, key, constants[key])


def _cell(value, ds):
    """SI coordinate -> grid index the way the analysers do it (NOT Solver.convert: can differ by one cell)."""
    return int(value / ds)


def _draw_field(plt, frame, limit, extent):
    image = plt.imshow(frame, vmax=limit, vmin=-limit, extent=extent, cmap='seismic')
    plt.colorbar(image)
    plt.ylabel("Y (m)")
    plt.xlabel("X (m)")
    return image


def _power_spectrum(samples, spacing, omega_max):
    """|FFT|^2 of the zero-padded series on 0 < omega < omega_max, omega in rad/s."""
    padded = np.zeros(FFT_POINTS)
    padded[:len(samples)] = samples
    omega = 2 * np.pi * np.fft.fftfreq(FFT_POINTS, d=spacing)
    band = (omega > 0) & (omega < omega_max)
    return omega[band], np.absolute(np.fft.fft(padded)[band]) ** 2


class FileLoader:

    def __init__(self, file_name):
        self.file_name = file_name
        self.constants = None
        self.h_matrix = None
        self.size = None
        self.load()

    def load(self):
        print("Loading npy file...")
        self.h_matrix = np.load(self.file_name + ".npy")
        try:
            handle = open(self.file_name + ".txt")
        except FileNotFoundError:
            print("Simulation constants txt file could not be found")
            return
        with handle:
            print("Loading txt file...")
            self.constants = json.load(handle)
        self.unpack_constants_from_json()

    def unpack_constants_from_json(self):
        _adopt(self, self.constants)
        self.material = np.array(self.constants['material'])
        self.size = len(self.material)
        self.pulses = self.constants['pulses']

    def _extent(self):
        return [0, self.length_x, self.length_y, 0]

    def play(self, interval=1, colourdepth=0.8, jupyter=False):
        plt, animation = _mpl()
        fig = plt.figure()
        limit = (1 - colourdepth) * np.max(self.h_matrix)
        image = _draw_field(plt, self.h_matrix[0], limit, self._extent())

        def show(k):
            image.set_array(self.h_matrix[k])
            fig.suptitle("Time step: {}".format(k * self.step_frequency))

        movie = animation.FuncAnimation(fig, show, frames=len(self.h_matrix), interval=interval, repeat=False)
        if jupyter:
            return movie
        plt.show()

    def plot_matrix_at_time(self, time):
        plt, _ = _mpl()
        plt.figure()
        frame = self.h_matrix[int(time / (self.dt * self.step_frequency))]
        _draw_field(plt, frame, np.max(self.h_matrix), self._extent())
        plt.suptitle("Time: " + str(time))
        plt.show()

    def get_matrix(self):
        return self.h_matrix

    def convert(self, value):
        return _cell(value, self.ds)

    def create_data_collector(self, location):
        return DataCollector(self.h_matrix, location, self.constants)


class DataCollector:
    """The time series of one probe cell of a snapshot stack, and its power spectrum."""

    def __init__(self, matrix, location, constants):
        self.matrix, self.location, self.constants = matrix, location, constants
        self.data = []
        self.fft_amplitude = self.fft_frequencies = None
        self.unpack_constants()
        self.collect_all()

    def unpack_constants(self):
        _adopt(self, self.constants)
        self.i_pos, self.j_pos = (self.convert(v) for v in self.location[:2])
        # one sample per SAVED snapshot, i.e. every step_frequency-th step
        self.time_vector = np.arange(0, len(self.matrix)) * self.dt * self.step_frequency

    def convert(self, value):
        return _cell(value, self.ds)

    def collect_all(self):
        self.data = self.matrix[:, self.i_pos, self.j_pos]

    def fft(self):
        self.fft_frequencies, self.fft_amplitude = _power_spectrum(self.data, self.dt * self.step_frequency,
                                                                   self.omega_max)

    def plot_time(self, show=True):
        plt, _ = _mpl()
        plt.suptitle("Detected pulse")
        plt.xlabel("Time (seconds)")
        plt.ylabel("Magnitude")
        plt.plot(self.time_vector, self.data)
        if show:
            plt.show()

    def plot_frequency(self, show=True, title=None):
        plt, _ = _mpl()
        if self.fft_amplitude is None:
            self.fft()
        if title is not None:
            plt.suptitle = title          # (sic) the reference assigns instead of calling; kept for identical behaviour
        plt.xlabel("Frequency (rad/s)")
        plt.ylabel("Amplitude")
        plt.plot(self.fft_frequencies, self.fft_amplitude)
        if show:
            plt.show()


class ProbeCollector(DataCollector):
    """NOT in the reference: a `DataCollector` whose series was sampled by an on-device probe during the run
    (`Solver.record_field`) instead of being sliced out of a saved snapshot stack -- same attributes, `fft()` and
    plots, no `.npy` needed."""

    def __init__(self, samples, location, constants):
        self._samples = np.asarray(samples, dtype=np.float64)
        super().__init__(self._samples, location, constants)        # len(matrix) = number of samples

    def collect_all(self):
        self.data = self._samples
