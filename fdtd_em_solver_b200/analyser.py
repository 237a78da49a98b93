"""Drop-in `analyser` module (reference analyser.py:7-159): consumers of the
`.npy` snapshot stack + `.txt` JSON constants that Solver.save/solve write.
Host-side numpy only -- out of the accelerated path; kept so
that reference scripts run unchanged against this package.
"""
import json

import numpy as np

_KEYS = ('dt', 'ds', 'length_x', 'length_y', 'stability', 'omega_max', 'end_time', 'step_frequency')


def _pyplot():
    import matplotlib.pyplot as plt
    return plt


class FileLoader:
    """reference analyser.py:7-84."""

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
            with open(self.file_name + ".txt") as json_file:
                print("Loading txt file...")
                self.constants = json.load(json_file)
        except FileNotFoundError:
            print("Simulation constants txt file could not be found")
            return
        self.unpack_constants_from_json()

    def unpack_constants_from_json(self):
        for key in _KEYS:
            setattr(self, key, self.constants[key])
        self.material = np.array(self.constants['material'])
        self.size = len(self.material)
        self.pulses = self.constants['pulses']

    def play(self, interval=1, colourdepth=0.8, jupyter=False):
        plt = _pyplot()
        import matplotlib.animation as animation
        scale = (1 - colourdepth) * np.max(self.h_matrix)
        fig, ax = plt.subplots()
        image = ax.imshow(self.h_matrix[0], vmax=scale, vmin=-scale, extent=[0, self.length_x, self.length_y, 0],
                          cmap='seismic')

        def animate(i):
            image.set_array(self.h_matrix[i])
            fig.suptitle("Time step: {}".format(i * self.step_frequency))

        plt.colorbar(image)
        plt.ylabel("Y (m)")
        plt.xlabel("X (m)")
        ani = animation.FuncAnimation(fig, animate, frames=len(self.h_matrix), interval=interval, repeat=False)
        if jupyter:
            return ani
        plt.show()

    def plot_matrix_at_time(self, time):
        plt = _pyplot()
        plt.figure()
        frame = int(time / (self.dt * self.step_frequency))
        top = np.max(self.h_matrix)
        im = plt.imshow(self.h_matrix[frame], vmax=top, vmin=-top, extent=[0, self.length_x, self.length_y, 0],
                        cmap='seismic')
        plt.colorbar(im)
        plt.suptitle("Time: " + str(time))
        plt.ylabel("Y (m)")
        plt.xlabel("X (m)")
        plt.show()

    def get_matrix(self):
        return self.h_matrix

    def convert(self, value):
        return int(value / self.ds)

    def create_data_collector(self, location):
        return DataCollector(self.h_matrix, location, self.constants)


class DataCollector:
    """reference analyser.py:87-159: one probe cell's time series and its power spectrum."""

    def __init__(self, matrix, location, constants):
        self.matrix = matrix
        self.location = location
        self.constants = constants
        self.data = []
        self.fft_amplitude = None
        self.fft_frequencies = None
        self.unpack_constants()
        self.collect_all()

    def unpack_constants(self):
        for key in _KEYS:
            setattr(self, key, self.constants[key])
        self.i_pos = self.convert(self.location[0])
        self.j_pos = self.convert(self.location[1])
        self.time_vector = np.arange(0, len(self.matrix)) * self.dt * self.step_frequency

    def convert(self, value):
        return int(value / self.ds)

    def collect_all(self):
        self.data = self.matrix[:, self.i_pos, self.j_pos]

    def plot_time(self, show=True):
        plt = _pyplot()
        plt.suptitle("Detected pulse")
        plt.xlabel("Time (seconds)")
        plt.ylabel("Magnitude")
        plt.plot(self.time_vector, self.data)
        if show:
            plt.show()

    def fft(self):
        """Zero-padded 2**14-point spectrum; sample spacing is dt*step_frequency (analyser.py:134-145)."""
        padded = np.concatenate((self.data, np.zeros(2 ** 14 - len(self.data))))
        freqs = 2 * np.pi * np.fft.fftfreq(len(padded), d=self.dt * self.step_frequency)
        keep = np.logical_and(freqs > 0, freqs < self.omega_max)
        self.fft_amplitude = np.absolute(np.fft.fft(padded)[keep]) ** 2
        self.fft_frequencies = freqs[keep]

    def plot_frequency(self, show=True, title=None):
        plt = _pyplot()
        if self.fft_amplitude is None:
            self.fft()
        if title is not None:
            plt.suptitle = title
        plt.xlabel("Frequency (rad/s)")
        plt.ylabel("Amplitude")
        plt.plot(self.fft_frequencies, self.fft_amplitude)
        if show:
            plt.show()
