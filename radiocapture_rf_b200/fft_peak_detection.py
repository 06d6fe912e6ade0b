"""Host-side peak picking of the scan path - mirrors fft_peak_detection.py:39-72 of the reference.

The reference loads the float32 vector fft_vector.py wrote (/tmp/fft_source_<i>, :39-40), shifts it by
|min| with an interpreted per-element loop (:58-59), takes the mean with Python's sum() over float32
scalars (:61), calls scipy.signal.find_peaks(width=[3 kHz, 30 kHz] in bins, prominence=1) (:65) and keeps
peaks above 2*mean (:71), mapping bin -> Hz with int(i*hz_per_bin - bw/2 + centre) (:72).  We call the
same scipy routine (it IS the reference's arithmetic) on the vector the GPU scan kernel produced; only
the O(L) interpreted loops are replaced by equivalent float32 numpy operations.
"""
import numpy as np


def load_vector(path):
    """numpy.fromfile(float32) exactly as fft_peak_detection.py:39-40."""
    with open(path, "rb") as fh:
        return np.fromfile(fh, np.float32)


def save_vector(path, vec):
    """The file fft_vector.py's blocks.file_sink writes: raw float32, one vector (fft_vector.py:44)."""
    np.asarray(vec, dtype=np.float32).tofile(path)


def detect_peaks(data, samp_rate, center_freq, fft_width=None):
    """Returns (bin indices, frequencies in Hz) of the detected carriers."""
    from scipy import signal
    data = np.array(data, dtype=np.float32, copy=True)
    if fft_width is None:
        fft_width = len(data)
    bandwidth = samp_rate
    hz_per_bin = bandwidth / fft_width
    min_width_in_bins = 3000 / hz_per_bin
    max_width_in_bins = 30000 / hz_per_bin
    data_min = data.min()
    data = (data + abs(data_min)).astype(np.float32)           # :58-59
    # :61  sum(data)/len(data): sequential float32 accumulation
    data_average = np.add.accumulate(data, dtype=np.float32)[-1] / np.float32(len(data))
    peaks = signal.find_peaks(data, width=[min_width_in_bins, max_width_in_bins], prominence=1)  # :65
    keep = [int(i) for i in peaks[0] if data[i] > data_average * 2]                               # :71
    freqs = [int((i * hz_per_bin) - (bandwidth / 2) + center_freq) for i in keep]                 # :72
    return np.asarray(keep, dtype=np.int64), np.asarray(freqs, dtype=np.int64)
