"""The opt-in single-pass pipeline (VELOSLAM_SINGLE_PASS=1: k_pose_pre + k_decode<.., FUSED>, the
segmentation scans run inside the decode kernel) against the same oracle comparisons as the default
two-pass pipeline: the parity tests of test_gpu_parity.py re-collected with the switch set."""
import numpy as np
import pytest

from veloslam_b200 import capi, synth

import parity as P
import test_gpu_parity as T

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _single_pass(monkeypatch):
    monkeypatch.setenv("VELOSLAM_SINGLE_PASS", "1")  # read by vs_create


def test_single_pass_is_engaged():
    """No k_scan / k_pose launch: the reset, the fused decode kernel and the frame-table gather."""
    pk, t = synth.hdl64_packets(600)
    ctx = P.make_ctx(synth.calib_hdl64())
    r = ctx.decode(synth.as_bytes(pk), np.ascontiguousarray(t), mode=capi.MODE_STREAMING,
                   carry=capi.carry_init(), t_base_us=int(t[0]))
    assert r.n_kernel_launches == 3          # k_reset, k_decode<FUSED>, k_frames
    ctx.close()


test_hdl64_decode_and_segmentation = T.test_hdl64_decode_and_segmentation
test_hdl64_identity_calibration_is_bit_exact = T.test_hdl64_identity_calibration_is_bit_exact
test_hdl32_stream_is_bit_exact = T.test_hdl32_stream_is_bit_exact
test_hdl32_azimuth_adjust_ties = T.test_hdl32_azimuth_adjust_ties
test_vlp16_mode = T.test_vlp16_mode
test_hdl64_deskew_against_ins_timeline = T.test_hdl64_deskew_against_ins_timeline
test_deskew_across_yaw_wrap_and_extrapolation = T.test_deskew_across_yaw_wrap_and_extrapolation
test_short_timeline_means_no_transform = T.test_short_timeline_means_no_transform
test_batches_with_carry_match_one_stream = T.test_batches_with_carry_match_one_stream
test_wrap_inside_packet_streaming_quirks = T.test_wrap_inside_packet_streaming_quirks
test_random_azimuths_streaming = T.test_random_azimuths_streaming
test_decreasing_azimuths_chain_of_nonconstant_maps = T.test_decreasing_azimuths_chain_of_nonconstant_maps
test_long_chain_of_nonconstant_skip_maps = T.test_long_chain_of_nonconstant_skip_maps
test_laser_selection_points_skip_and_crop = T.test_laser_selection_points_skip_and_crop
test_ragged_sizes = T.test_ragged_sizes
test_sizes_around_the_decode_tile = T.test_sizes_around_the_decode_tile
test_sparse_returns_partial_granules = T.test_sparse_returns_partial_granules
test_odd_and_even_halos = T.test_odd_and_even_halos
test_offline_mode_matches_get_frame = T.test_offline_mode_matches_get_frame
test_offline_random_azimuths = T.test_offline_random_azimuths
test_pcap_record_stride_and_gpu_side_times = T.test_pcap_record_stride_and_gpu_side_times
test_halo_shard_matches_whole_stream = T.test_halo_shard_matches_whole_stream
