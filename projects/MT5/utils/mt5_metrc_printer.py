"""Console printer that also reports samples/s of the encoder-decoder job (reference
projects/MT5/utils/mt5_metrc_printer.py); the default ``CommonMetricPrinter`` already prints throughput and
tokens/s, this subclass only adds the decoder token rate."""
from libai_b200.utils.events import CommonMetricPrinter, get_event_storage


class MT5MetricPrinter(CommonMetricPrinter):
    def __init__(self, batch_size, max_iter, log_period, decoder_seq_length=128):
        super().__init__(batch_size, max_iter, log_period)
        self.decoder_seq_length = decoder_seq_length

    def write(self):
        super().write()
        storage = get_event_storage()
        try:
            t = storage.history("time").global_avg()
            if t > 0:
                self.logger.info(f" decoder tokens/s: {self.batch_size * self.decoder_seq_length / t:.0f}")
        except KeyError:
            pass
