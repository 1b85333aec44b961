"""The committed ncu launch lists under profiles/ stay readable by the tools that summarise them (CPU)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def test_launch_lists_with_dram_bytes_parse():
    import launch_bytes
    text = launch_bytes.summarise(os.path.join(ROOT, 'profiles', 'r02x_launches_stage1_bytes.csv'))
    lines = text.splitlines()
    assert lines[0].startswith('total') and 'umma_chain_kernel' in text and 'umma_wgrad_mn_kernel' in text
    # the chain kernels dominate the step and the tangent sweep moves > 4 TB/s of DRAM traffic
    chain2 = next(l for l in lines if 'umma_chain_kernel<2>' in l).split()
    assert float(chain2[6]) > 4000          # columns: us, share, n, avg us, rd MB, wr MB, GB/s, kernel


def test_chain_traffic_record_matches_its_source():
    import csv
    import json
    tj = json.load(open(os.path.join(ROOT, 'profiles', 'chain_traffic.json')))
    rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', 'r02h_chain_reverse_sweep_raw.csv'))))
    hdr, units, val = rows[0], rows[1], rows[2]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    rd = float(val[hdr.index('dram__bytes_read.sum')]) * scale[units[hdr.index('dram__bytes_read.sum')]]
    wr = float(val[hdr.index('dram__bytes_write.sum')]) * scale[units[hdr.index('dram__bytes_write.sum')]]
    assert abs((rd + wr) - tj['dram_bytes_per_launch']) <= 1e-3 * tj['dram_bytes_per_launch']
    assert abs(tj['dram_bytes_per_row'] * tj['rows'] - tj['dram_bytes_per_launch']) <= 1e-3 * tj['dram_bytes_per_launch']
