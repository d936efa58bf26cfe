"""TEST INFRASTRUCTURE (oracle) - cycle-level emulation of the reference's fixed-point generator RTL.

Not part of the product path: only tests/ may import this.  It exists to pin the closed-form
`rtl_literal` semantics (oracle/ofdmgan_oracle.c, ofdmgan_gen_fwd_q mode 1) to what the committed
Verilog actually computes, since no Verilog simulator is available.

Follows, block by block (nonblocking-assignment semantics: every register update is computed from the
pre-edge values and committed together):
  rtl/ofdmGAN/generator_mini.v:165-174   state register
  rtl/ofdmGAN/generator_mini.v:176-234   next-state logic
  rtl/ofdmGAN/generator_mini.v:239-276   input loading / output index counters
  rtl/ofdmGAN/generator_mini.v:283-649   pipelined convolution processing (3-stage pipe, accumulators)
  rtl/ofdmGAN/generator_mini.v:655-666   output register
  rtl/ofdmGAN/generator_mini.v:119-146   ROM ports, parallel multipliers, per-tap `>>> 7`
  rtl/ofdmGAN/weight_rom.v:164-166,261-263   synchronous ROM reads (one-cycle latency)

Pinned against tests/golden/rtl_generator_vectors.json (10 known-answer frames + accumulate-stage traces
extracted from rtl/ofdmGAN/tb_generator_mini.vcd) by tests/test_oracle_fixed_point.py.
"""

(ST_IDLE, ST_LOAD_IN, ST_ENC1, ST_BNECK, ST_UPSAMPLE1, ST_DEC1, ST_SKIP_ADD, ST_UPSAMPLE2,
 ST_OUT_CONV, ST_TANH, ST_OUTPUT, ST_DONE) = range(12)

IN_CH, OUT_CH, FRAME_LEN = 2, 2, 16
ENC1_OUT_CH, ENC1_OUT_LEN = 4, 8
BNECK_OUT_CH, BNECK_OUT_LEN = 8, 4
DEC1_OUT_CH, DEC1_OUT_LEN = 4, 8
UP1_LEN = 8
WADDR_ENC1, WADDR_BNECK, WADDR_DEC1, WADDR_OUT = 0, 24, 120, 216
BADDR_ENC1, BADDR_BNECK, BADDR_DEC1, BADDR_OUT = 0, 4, 12, 16


def _s(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def _sat16(v):
    return 32767 if v > 32767 else (-32768 if v < -32768 else v)


def _lrelu16(r):
    # generator_mini.v:359-360 : 16-bit context, arithmetic shifts, wrapping add
    return _s((r >> 2) + (r >> 4), 16) if r < 0 else r


class _Buf2D:
    """reg array [rows][cols]; out-of-range writes are dropped, out-of-range reads give 0 (Verilog x)."""

    def __init__(self, rows, cols):
        self.rows, self.cols = rows, cols
        self.d = [[0] * cols for _ in range(rows)]

    def rd(self, r, c):
        if 0 <= r < self.rows and 0 <= c < self.cols:
            return self.d[r][c]
        return 0

    def wr(self, r, c, v):
        if 0 <= r < self.rows and 0 <= c < self.cols:
            self.d[r][c] = v


class GeneratorMiniRTL:
    def __init__(self, weights, biases):
        """weights: dict/list addr->int8 (2048 deep), biases: addr->int16 (64 deep)."""
        self.W = [0] * 2048
        self.B = [0] * 64
        for k, v in (weights.items() if isinstance(weights, dict) else enumerate(weights)):
            self.W[int(k)] = int(v)
        for k, v in (biases.items() if isinstance(biases, dict) else enumerate(biases)):
            self.B[int(k)] = int(v)
        self.reset()

    def reset(self):
        self.state = ST_IDLE
        self.in_ch_cnt = self.in_pos_cnt = 0
        self.out_ch_cnt = self.out_pos_cnt = self.in_ch_iter = self.pipe_flush = 0
        self.weight_addr_base = 0
        self.bias_addr = 0
        self.data_k = [0, 0, 0]
        self.weight_k = [0, 0, 0]      # ROM output registers (weight_rom.v:164-166); x after reset -> 0
        self.bias_data = 0
        self.s2 = dict(valid=0, out_ch=0, out_pos=0, last=0)
        self.s3 = dict(valid=0, out_ch=0, out_pos=0, last=0, ksum=0)
        self.accum = [0] * 16
        self.input_buf = _Buf2D(IN_CH, FRAME_LEN + 2)
        self.skip_buf = _Buf2D(ENC1_OUT_CH, ENC1_OUT_LEN)
        self.enc1_buf = _Buf2D(ENC1_OUT_CH, ENC1_OUT_LEN + 2)
        self.bneck_buf = _Buf2D(BNECK_OUT_CH, BNECK_OUT_LEN)
        self.up1_buf = _Buf2D(BNECK_OUT_CH, UP1_LEN + 2)
        self.dec1_buf = _Buf2D(DEC1_OUT_CH, DEC1_OUT_LEN)
        self.up2_buf = _Buf2D(DEC1_OUT_CH, FRAME_LEN)
        self.out_buf = _Buf2D(OUT_CH, FRAME_LEN)
        self.data_out = 0
        self.valid_out = 0
        self.trace = []

    # combinational helpers -------------------------------------------------------------------
    def _mults(self):
        return [self.data_k[i] * self.weight_k[i] for i in range(3)]

    def _kernel_sum(self):
        m = self._mults()
        return _s((m[0] >> 7) + (m[1] >> 7) + (m[2] >> 7), 32)

    def _next_state(self, start, valid_in, ready_out):
        s, oc, op, it, fl = self.state, self.out_ch_cnt, self.out_pos_cnt, self.in_ch_iter, self.pipe_flush
        if s == ST_IDLE:
            return ST_LOAD_IN if start else s
        if s == ST_LOAD_IN:
            return ST_ENC1 if (self.in_ch_cnt == IN_CH - 1 and self.in_pos_cnt == FRAME_LEN - 1 and valid_in) else s
        if s == ST_ENC1:
            return ST_BNECK if (oc == ENC1_OUT_CH - 1 and op == ENC1_OUT_LEN - 1 and it == IN_CH - 1 and fl == 2) else s
        if s == ST_BNECK:
            return ST_UPSAMPLE1 if (oc == BNECK_OUT_CH - 1 and op == BNECK_OUT_LEN - 1 and it == ENC1_OUT_CH - 1 and fl == 2) else s
        if s == ST_UPSAMPLE1:
            return ST_DEC1 if (oc == BNECK_OUT_CH - 1 and op == BNECK_OUT_LEN - 1) else s
        if s == ST_DEC1:
            return ST_SKIP_ADD if (oc == DEC1_OUT_CH - 1 and op == DEC1_OUT_LEN - 1 and it == BNECK_OUT_CH - 1 and fl == 2) else s
        if s == ST_SKIP_ADD:
            return ST_UPSAMPLE2 if (oc == DEC1_OUT_CH - 1 and op == DEC1_OUT_LEN - 1) else s
        if s == ST_UPSAMPLE2:
            return ST_OUT_CONV if (oc == DEC1_OUT_CH - 1 and op == DEC1_OUT_LEN - 1) else s
        if s == ST_OUT_CONV:
            return ST_TANH if (oc == OUT_CH - 1 and op == FRAME_LEN - 1 and it == DEC1_OUT_CH - 1 and fl == 2) else s
        if s == ST_TANH:
            return ST_OUTPUT if (oc == OUT_CH - 1 and op == FRAME_LEN - 1) else s
        if s == ST_OUTPUT:
            return ST_DONE if (self.in_ch_cnt == OUT_CH - 1 and self.in_pos_cnt == FRAME_LEN - 1 and ready_out) else s
        return ST_IDLE

    # one rising clock edge -------------------------------------------------------------------
    def clock(self, start=0, data_in=0, valid_in=0, ready_out=1):
        st = self.state
        nxt = self._next_state(start, valid_in, ready_out)
        upd = {}            # scalar register updates
        wr = []             # (buffer, r, c, v)
        acc_upd = {}

        # ROM synchronous reads (addresses are 11 / 6 bit wide)
        new_weight_k = [self.W[(self.weight_addr_base + i) & 0x7FF] for i in range(3)]
        new_bias = self.B[self.bias_addr & 0x3F]

        # ---- input loading block (generator_mini.v:239-276)
        if st == ST_IDLE and start:
            upd["in_ch_cnt"] = 0
            upd["in_pos_cnt"] = 0
            for r in range(IN_CH):
                for c in range(FRAME_LEN + 2):
                    wr.append((self.input_buf, r, c, 0))
        elif st == ST_LOAD_IN and valid_in:
            wr.append((self.input_buf, self.in_ch_cnt, self.in_pos_cnt + 1, _s(data_in, 16)))
            if self.in_pos_cnt == FRAME_LEN - 1:
                upd["in_pos_cnt"] = 0
                upd["in_ch_cnt"] = (self.in_ch_cnt + 1) & 7
            else:
                upd["in_pos_cnt"] = (self.in_pos_cnt + 1) & 31
        elif st == ST_OUTPUT and ready_out:
            if self.in_pos_cnt == FRAME_LEN - 1:
                upd["in_pos_cnt"] = 0
                upd["in_ch_cnt"] = (self.in_ch_cnt + 1) & 7
            else:
                upd["in_pos_cnt"] = (self.in_pos_cnt + 1) & 31

        # ---- processing block (generator_mini.v:283-649)
        oc, op, it, fl = self.out_ch_cnt, self.out_pos_cnt, self.in_ch_iter, self.pipe_flush
        s2, s3 = self.s2, self.s3
        new_s2, new_s3 = dict(s2), dict(s3)

        def conv_stage(src, stride, in_ch, out_ch, out_len, waddr, baddr, wmul, store, k1=False):
            upd["weight_addr_base"] = (waddr + oc * wmul + it * (1 if k1 else 3)) & 0x7FF
            upd["bias_addr"] = (baddr + oc) & 0x3F
            if k1:
                upd["data_k"] = [src.rd(it, op), self.data_k[1], self.data_k[2]]
            else:
                upd["data_k"] = [src.rd(it, op * stride + k) for k in range(3)]
            new_s2.update(valid=1, out_ch=oc, out_pos=op, last=int(it == in_ch - 1))
            ksum = _s(self._mults()[0] >> 7, 32) if k1 else self._kernel_sum()
            new_s3.update(valid=s2["valid"], out_ch=s2["out_ch"], out_pos=s2["out_pos"], last=s2["last"], ksum=ksum)
            if s3["valid"]:
                self.trace.append([st, s3["out_ch"], s3["out_pos"], s3["last"], s3["ksum"]])
                if s3["last"]:
                    total = _s(self.accum[s3["out_ch"]] + s3["ksum"] + self.bias_data, 32)
                    store(s3["out_ch"], s3["out_pos"], _sat16(total))
                    acc_upd[s3["out_ch"]] = 0
                else:
                    acc_upd[s3["out_ch"]] = _s(self.accum[s3["out_ch"]] + s3["ksum"], 32)
            if it == in_ch - 1:
                upd["in_ch_iter"] = 0
                if op == out_len - 1:
                    upd["out_pos_cnt"] = 0
                    if oc == out_ch - 1:
                        upd["pipe_flush"] = (fl + 1) & 7
                    else:
                        upd["out_ch_cnt"] = (oc + 1) & 15
                else:
                    upd["out_pos_cnt"] = (op + 1) & 31
            else:
                upd["in_ch_iter"] = (it + 1) & 15

        def clear_pipe_and_acc():
            new_s2["valid"] = 0
            new_s3["valid"] = 0
            for i in range(16):
                acc_upd[i] = 0

        def step_pos(out_len, out_ch):
            if op == out_len - 1:
                upd["out_pos_cnt"] = 0
                upd["out_ch_cnt"] = 0 if oc == out_ch - 1 else (oc + 1) & 15
            else:
                upd["out_pos_cnt"] = (op + 1) & 31

        if st in (ST_IDLE, ST_LOAD_IN):
            upd.update(out_ch_cnt=0, out_pos_cnt=0, in_ch_iter=0, pipe_flush=0)
            clear_pipe_and_acc()
        elif st == ST_ENC1:
            def store(c, p, v):
                v = _lrelu16(v)
                wr.append((self.enc1_buf, c, p + 1, v))
                wr.append((self.skip_buf, c, p, v))
            conv_stage(self.input_buf, 2, IN_CH, ENC1_OUT_CH, ENC1_OUT_LEN, WADDR_ENC1, BADDR_ENC1, IN_CH * 3, store)
        elif st == ST_BNECK:
            if oc == 0 and op == 0 and it == 0 and fl == 0:
                # generator_mini.v:390-393; later nonblocking assignments in the same block override s2/s3 valid
                for i in range(16):
                    acc_upd[i] = 0
            def store(c, p, v):
                wr.append((self.bneck_buf, c, p, _lrelu16(v)))
            conv_stage(self.enc1_buf, 2, ENC1_OUT_CH, BNECK_OUT_CH, BNECK_OUT_LEN, WADDR_BNECK, BADDR_BNECK,
                       ENC1_OUT_CH * 3, store)
        elif st == ST_UPSAMPLE1:
            clear_pipe_and_acc()
            upd["pipe_flush"] = 0
            v = self.bneck_buf.rd(oc, op)
            wr.append((self.up1_buf, oc, op * 2 + 1, v))
            wr.append((self.up1_buf, oc, op * 2 + 2, v))
            step_pos(BNECK_OUT_LEN, BNECK_OUT_CH)
        elif st == ST_DEC1:
            def store(c, p, v):
                wr.append((self.dec1_buf, c, p, _lrelu16(v)))
            conv_stage(self.up1_buf, 1, BNECK_OUT_CH, DEC1_OUT_CH, DEC1_OUT_LEN, WADDR_DEC1, BADDR_DEC1,
                       BNECK_OUT_CH * 3, store)
        elif st == ST_SKIP_ADD:
            clear_pipe_and_acc()
            upd["pipe_flush"] = 0
            wr.append((self.dec1_buf, oc, op, _sat16(self.dec1_buf.rd(oc, op) + self.skip_buf.rd(oc, op))))
            step_pos(DEC1_OUT_LEN, DEC1_OUT_CH)
        elif st == ST_UPSAMPLE2:
            v = self.dec1_buf.rd(oc, op)
            wr.append((self.up2_buf, oc, op * 2, v))
            wr.append((self.up2_buf, oc, op * 2 + 1, v))
            step_pos(DEC1_OUT_LEN, DEC1_OUT_CH)
        elif st == ST_OUT_CONV:
            def store(c, p, v):
                wr.append((self.out_buf, c, p, v))
            conv_stage(self.up2_buf, 1, DEC1_OUT_CH, OUT_CH, FRAME_LEN, WADDR_OUT, BADDR_OUT, DEC1_OUT_CH, store, k1=True)
        elif st == ST_TANH:
            new_s2["valid"] = 0
            new_s3["valid"] = 0
            upd["pipe_flush"] = 0
            v = self.out_buf.rd(oc, op)
            if v > 0x0100:
                wr.append((self.out_buf, oc, op, 0x00FF))
            elif v < -0x0100:
                wr.append((self.out_buf, oc, op, _s(0xFF01, 16)))
            if op == FRAME_LEN - 1:
                upd["out_pos_cnt"] = 0
                if oc == OUT_CH - 1:
                    upd["in_ch_cnt"] = 0
                    upd["in_pos_cnt"] = 0
                upd["out_ch_cnt"] = (oc + 1) & 15
            else:
                upd["out_pos_cnt"] = (op + 1) & 31

        # ---- output register block (generator_mini.v:655-666)
        if st == ST_OUTPUT:
            new_data_out = self.out_buf.rd(self.in_ch_cnt, self.in_pos_cnt)
            new_valid_out = 1
        else:
            new_data_out = self.data_out
            new_valid_out = 0

        # ---- commit
        for b, r, c, v in wr:
            b.wr(r, c, v)
        for k, v in acc_upd.items():
            self.accum[k] = v
        for k, v in upd.items():
            setattr(self, k, v)
        self.s2, self.s3 = new_s2, new_s3
        self.weight_k, self.bias_data = new_weight_k, new_bias
        self.data_out, self.valid_out = new_data_out, new_valid_out
        self.state = nxt

    # frame-level driver ----------------------------------------------------------------------
    def run_frame(self, frame32, max_cycles=5000):
        """Drive one frame the way tb_generator_mini.v:run_test does (start pulse, 32 valid samples, ready_out=1).

        Returns (output32, cycles from start to done)."""
        assert len(frame32) == 32
        for _ in range(3):
            self.clock()
        self.clock(start=1)
        cycles = 1
        idx = 0
        out = []
        while self.state != ST_DONE and cycles < max_cycles:
            if self.state == ST_LOAD_IN and idx < 32:
                self.clock(data_in=frame32[idx], valid_in=1)
                idx += 1
            else:
                self.clock()
            cycles += 1
            if self.valid_out and len(out) < 32:
                out.append(self.data_out)
        self.clock()
        return out, cycles


# =====================================================================================================================
# Fixed-point critic: cycle-level emulation of rtl/ofdmGAN/discriminator_mini.v
#   :165-211  state register / next-state logic        :216-256  candidate / condition loading
#   :261-479  pipelined conv1, conv2, sum-pool, dense   :484-500  score register (saturated dense accumulator)
#   :104-134  three weight-ROM ports + bias ROM (synchronous, one-cycle latency), per-tap `>>> 7`
# Pinned against tests/golden/rtl_critic_vectors.json (5 scores + every accumulate-stage entry of the committed Icarus
# run rtl/ofdmGAN/tb_discriminator_mini.vcd) by tests/test_oracle_fixed_point.py.
# =====================================================================================================================
(DS_IDLE, DS_LOAD_CAND, DS_LOAD_COND, DS_CONV1, DS_CONV2, DS_POOL, DS_DENSE, DS_OUTPUT, DS_DONE) = range(9)
D_IN_CH, D_C1_CH, D_C1_LEN, D_C2_CH, D_C2_LEN = 4, 8, 8, 16, 4
D_WADDR_C1, D_WADDR_C2, D_WADDR_DENSE = 256, 352, 736
D_BADDR_C1, D_BADDR_C2, D_BADDR_DENSE = 32, 40, 56


class DiscriminatorMiniRTL:
    def __init__(self, weights, biases):
        self.W = [0] * 2048
        self.B = [0] * 64
        for k, v in (weights.items() if isinstance(weights, dict) else enumerate(weights)):
            self.W[int(k)] = int(v)
        for k, v in (biases.items() if isinstance(biases, dict) else enumerate(biases)):
            self.B[int(k)] = int(v)
        self.reset()

    def reset(self):
        self.state = DS_IDLE
        self.load_ch_cnt = self.load_pos_cnt = 0
        self.out_ch_cnt = self.out_pos_cnt = self.in_ch_iter = self.pipe_flush = 0
        self.weight_addr_base = self.bias_addr = 0
        self.dense_acc = 0
        self.data_k = [0, 0, 0]
        self.weight_k = [0, 0, 0]
        self.bias_data = 0
        self.s2 = dict(valid=0, out_ch=0, out_pos=0, last=0)
        self.s3 = dict(valid=0, out_ch=0, out_pos=0, last=0, ksum=0)
        self.accum = [0] * 16
        self.input_buf = _Buf2D(D_IN_CH, FRAME_LEN + 2)
        self.conv1_buf = _Buf2D(D_C1_CH, D_C1_LEN + 2)
        self.conv2_buf = _Buf2D(D_C2_CH, D_C2_LEN)
        self.pool_buf = [0] * D_C2_CH
        self.score_out, self.score_valid = 0, 0
        self.trace = []

    def _next_state(self, start, cand_valid, cond_valid):
        s, oc, op, it, fl = self.state, self.out_ch_cnt, self.out_pos_cnt, self.in_ch_iter, self.pipe_flush
        full = self.load_ch_cnt == 1 and self.load_pos_cnt == FRAME_LEN - 1
        if s == DS_IDLE:
            return DS_LOAD_CAND if start else s
        if s == DS_LOAD_CAND:
            return DS_LOAD_COND if full and cand_valid else s
        if s == DS_LOAD_COND:
            return DS_CONV1 if full and cond_valid else s
        if s == DS_CONV1:
            return DS_CONV2 if (oc == D_C1_CH - 1 and op == D_C1_LEN - 1 and it == D_IN_CH - 1 and fl == 2) else s
        if s == DS_CONV2:
            return DS_POOL if (oc == D_C2_CH - 1 and op == D_C2_LEN - 1 and it == D_C1_CH - 1 and fl == 2) else s
        if s == DS_POOL:
            return DS_DENSE if (oc == D_C2_CH - 1 and op == D_C2_LEN - 1) else s
        if s == DS_DENSE:
            return DS_OUTPUT if (oc == D_C2_CH - 1 and fl == 2) else s
        if s == DS_OUTPUT:
            return DS_DONE
        return DS_IDLE

    def clock(self, start=0, cand_in=0, cand_valid=0, cond_in=0, cond_valid=0):
        st = self.state
        nxt = self._next_state(start, cand_valid, cond_valid)
        upd, wr, acc_upd, pool_upd = {}, [], {}, {}
        new_weight_k = [self.W[(self.weight_addr_base + i) & 0x7FF] for i in range(3)]
        new_bias = self.B[self.bias_addr & 0x3F]

        # ---- input loading (discriminator_mini.v:216-256)
        if st == DS_IDLE and start:
            upd["load_ch_cnt"] = upd["load_pos_cnt"] = 0
            for r in range(D_IN_CH):
                for c in range(FRAME_LEN + 2):
                    wr.append((self.input_buf, r, c, 0))
        elif st == DS_LOAD_CAND and cand_valid:
            wr.append((self.input_buf, self.load_ch_cnt, self.load_pos_cnt + 1, _s(cand_in, 16)))
            if self.load_pos_cnt == FRAME_LEN - 1:
                upd["load_pos_cnt"] = 0
                upd["load_ch_cnt"] = 0 if self.load_ch_cnt == 1 else (self.load_ch_cnt + 1) & 3
            else:
                upd["load_pos_cnt"] = (self.load_pos_cnt + 1) & 31
        elif st == DS_LOAD_COND and cond_valid:
            wr.append((self.input_buf, self.load_ch_cnt + 2, self.load_pos_cnt + 1, _s(cond_in, 16)))
            if self.load_pos_cnt == FRAME_LEN - 1:
                upd["load_pos_cnt"] = 0
                upd["load_ch_cnt"] = (self.load_ch_cnt + 1) & 3
            else:
                upd["load_pos_cnt"] = (self.load_pos_cnt + 1) & 31

        # ---- processing (discriminator_mini.v:261-479)
        oc, op, it, fl = self.out_ch_cnt, self.out_pos_cnt, self.in_ch_iter, self.pipe_flush
        s2, s3 = self.s2, self.s3
        new_s2, new_s3 = dict(s2), dict(s3)
        mults = [self.data_k[i] * self.weight_k[i] for i in range(3)]
        kernel_sum = _s((mults[0] >> 7) + (mults[1] >> 7) + (mults[2] >> 7), 32)

        def conv_stage(src, in_ch, out_ch, out_len, waddr, baddr, store):
            upd["weight_addr_base"] = (waddr + oc * (in_ch * 3) + it * 3) & 0x7FF
            upd["bias_addr"] = (baddr + oc) & 0x3F
            upd["data_k"] = [src.rd(it, op * 2 + k) for k in range(3)]
            new_s2.update(valid=1, out_ch=oc, out_pos=op, last=int(it == in_ch - 1))
            new_s3.update(valid=s2["valid"], out_ch=s2["out_ch"], out_pos=s2["out_pos"], last=s2["last"], ksum=kernel_sum)
            if s3["valid"]:
                self.trace.append([st, s3["out_ch"], s3["out_pos"], s3["last"], s3["ksum"]])
                a = s3["out_ch"] & 15
                if s3["last"]:
                    store(s3["out_ch"], s3["out_pos"], _lrelu16(_sat16(_s(self.accum[a] + s3["ksum"] + self.bias_data, 32))))
                    acc_upd[a] = 0
                else:
                    acc_upd[a] = _s(self.accum[a] + s3["ksum"], 32)
            if it == in_ch - 1:
                upd["in_ch_iter"] = 0
                if op == out_len - 1:
                    upd["out_pos_cnt"] = 0
                    if oc == out_ch - 1:
                        upd["pipe_flush"] = (fl + 1) & 7
                    else:
                        upd["out_ch_cnt"] = (oc + 1) & 31
                else:
                    upd["out_pos_cnt"] = (op + 1) & 31
            else:
                upd["in_ch_iter"] = (it + 1) & 31

        if st in (DS_IDLE, DS_LOAD_CAND, DS_LOAD_COND):
            upd.update(out_ch_cnt=0, out_pos_cnt=0, in_ch_iter=0, pipe_flush=0, dense_acc=0)
            new_s2["valid"] = new_s3["valid"] = 0
            for i in range(16):
                acc_upd[i] = 0
                pool_upd[i] = 0
        elif st == DS_CONV1:
            conv_stage(self.input_buf, D_IN_CH, D_C1_CH, D_C1_LEN, D_WADDR_C1, D_BADDR_C1,
                       lambda c, p, v: wr.append((self.conv1_buf, c, p + 1, v)))
        elif st == DS_CONV2:
            if oc == 0 and op == 0 and it == 0 and fl == 0:            # :359-362 (the later assignments override valid)
                for i in range(16):
                    acc_upd[i] = 0
            conv_stage(self.conv1_buf, D_C1_CH, D_C2_CH, D_C2_LEN, D_WADDR_C2, D_BADDR_C2,
                       lambda c, p, v: wr.append((self.conv2_buf, c, p, v)))
        elif st == DS_POOL:
            new_s2["valid"] = new_s3["valid"] = 0
            upd["pipe_flush"] = 0
            pool_upd[oc & 15] = _s(self.pool_buf[oc & 15] + self.conv2_buf.rd(oc, op), 32)
            if op == D_C2_LEN - 1:
                upd["out_pos_cnt"] = 0
                upd["out_ch_cnt"] = 0 if oc == D_C2_CH - 1 else (oc + 1) & 31
            else:
                upd["out_pos_cnt"] = (op + 1) & 31
        elif st == DS_DENSE:
            upd["weight_addr_base"] = (D_WADDR_DENSE + oc) & 0x7FF
            upd["bias_addr"] = D_BADDR_DENSE
            upd["data_k"] = [_s(self.pool_buf[oc & 15], 16), self.data_k[1], self.data_k[2]]
            new_s2.update(valid=1, out_ch=oc, last=int(oc == D_C2_CH - 1))
            new_s3.update(valid=s2["valid"], out_ch=s2["out_ch"], last=s2["last"], ksum=_s(mults[0] >> 7, 32))
            if s3["valid"]:
                self.trace.append([st, s3["out_ch"], s3["out_pos"], s3["last"], s3["ksum"]])
                upd["dense_acc"] = _s(self.dense_acc + s3["ksum"] + (self.bias_data if s3["last"] else 0), 32)
            if oc == D_C2_CH - 1:
                upd["pipe_flush"] = (fl + 1) & 7
            else:
                upd["out_ch_cnt"] = (oc + 1) & 31

        # ---- score register (:484-500)
        if st == DS_OUTPUT:
            new_score, new_valid = _sat16(self.dense_acc), 1
        else:
            new_score, new_valid = self.score_out, 0

        for b, r, c, v in wr:
            b.wr(r, c, v)
        for k, v in acc_upd.items():
            self.accum[k] = v
        for k, v in pool_upd.items():
            self.pool_buf[k] = v
        for k, v in upd.items():
            setattr(self, k, v)
        self.s2, self.s3 = new_s2, new_s3
        self.weight_k, self.bias_data = new_weight_k, new_bias
        self.score_out, self.score_valid = new_score, new_valid
        self.state = nxt

    def run_frame(self, cand32, cond32, max_cycles=5000):
        """Drive one (candidate, condition) pair as tb_discriminator_mini.v:run_test does.  Returns (score, cycles)."""
        assert len(cand32) == 32 and len(cond32) == 32
        for _ in range(3):
            self.clock()
        self.clock(start=1)
        cycles, ia, ib, score = 1, 0, 0, None
        while self.state != DS_DONE and cycles < max_cycles:
            if self.state == DS_LOAD_CAND and ia < 32:
                self.clock(cand_in=cand32[ia], cand_valid=1)
                ia += 1
            elif self.state == DS_LOAD_COND and ib < 32:
                self.clock(cond_in=cond32[ib], cond_valid=1)
                ib += 1
            else:
                self.clock()
            cycles += 1
            if self.score_valid:
                score = self.score_out
        self.clock()
        return score, cycles
