#!/bin/bash
for flags in "" "" "--no-clocks" "--no-clocks" "--steps 10"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --skip-extra --cpu-seconds 1 $flags > /tmp/b.json 2>/tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('[$flags]', '%.3e'%d['value'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'e2e %.3e'%d['e2e']['value'], d['step_ms'], d['clocks'])"
done
