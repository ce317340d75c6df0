// Dedicated Montgomery squaring by product scanning -- an experiment, not product code.
// Paste into struct Mont<C> of masp_b200/csrc/field.cuh (it uses its carry-chain helpers);
// sqr_ps_test.cpp checks it against Mont::mul_portable on the host model of the PTX carry flag.
// Measured on B200 inside msm_accumulate_g1 (2 of the 10 multiplications of a mixed addition are
// squarings): 166 registers either way, 6828 instead of 7158 IMAD.WIDE in the kernel, but
// 480 proofs/s against 490 and 25.1 ms against 24.3 ms per accumulation launch.
    // a^2 * 2^(-32N) mod p with the symmetry of the square: each cross product a_i a_j (i < j) is
    // formed once (N(N-1)/2 wide multiplies instead of N^2 - N), the sum is doubled by a one-bit
    // shift, the N diagonal squares join in one carry chain; the Montgomery reduction then runs in
    // product-scanning order over the 2N-limb square.  Column sums live in a three-word accumulator,
    // so every wide multiply carries one extra add-with-carry on the (idle) ALU pipe.  N^2 + N(N+1)/2
    // wide multiplies instead of 2 N^2: 222 against 288 for Fp.
    MB_HD static void acc3(uint32_t x, uint32_t y, uint32_t& c0, uint32_t& c1, uint32_t& c2) {
        c0 = mad_lo_cc(x, y, c0);
        c1 = madc_hi_cc(x, y, c1);
        c2 = addc(c2, 0);
    }
    MB_HD static Mont sqr_inline(const Mont& a) {
        uint32_t t[2 * N];
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        t[0] = 0;
        MB_UNROLL
        for (int k = 1; k <= 2 * N - 3; ++k) {
            MB_UNROLL
            for (int i = (k >= N ? k - N + 1 : 0); 2 * i < k; ++i) acc3(a.v[i], a.v[k - i], c0, c1, c2);
            t[k] = c0;
            c0 = c1;
            c1 = c2;
            c2 = 0;
        }
        t[2 * N - 2] = c0;
        t[2 * N - 1] = c1;
        MB_UNROLL
        for (int k = 2 * N - 1; k >= 1; --k) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
        t[0] = mad_lo_cc(a.v[0], a.v[0], 0);
        t[1] = madc_hi_cc(a.v[0], a.v[0], t[1]);
        MB_UNROLL
        for (int i = 1; i < N; ++i) {
            t[2 * i] = madc_lo_cc(a.v[i], a.v[i], t[2 * i]);
            t[2 * i + 1] = madc_hi_cc(a.v[i], a.v[i], t[2 * i + 1]);
        }
        // Montgomery reduction, column by column
        uint32_t m[N];
        c0 = c1 = c2 = 0;
        MB_UNROLL
        for (int k = 0; k < N; ++k) {
            MB_UNROLL
            for (int i = 0; i < k; ++i) acc3(m[i], C::mod(k - i), c0, c1, c2);
            c0 = add_cc(c0, t[k]);
            c1 = addc_cc(c1, 0);
            c2 = addc(c2, 0);
            m[k] = mul_lo(c0, C::INV);
            acc3(m[k], C::mod(0), c0, c1, c2);
            c0 = c1;
            c1 = c2;
            c2 = 0;
        }
        Mont s, u;
        MB_UNROLL
        for (int k = N; k < 2 * N; ++k) {
            MB_UNROLL
            for (int i = k - N + 1; i < N; ++i) acc3(m[i], C::mod(k - i), c0, c1, c2);
            c0 = add_cc(c0, t[k]);
            c1 = addc_cc(c1, 0);
            c2 = addc(c2, 0);
            s.v[k - N] = c0;
            c0 = c1;
            c1 = c2;
            c2 = 0;
        }
        // s < 2p (a < p): one conditional subtraction
        u.v[0] = sub_cc(s.v[0], C::mod(0));
        MB_UNROLL
        for (int i = 1; i < N; ++i) u.v[i] = subc_cc(s.v[i], C::mod(i));
        uint32_t borrow = subc(0, 0);
        MB_UNROLL
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : u.v[i];
        return s;
    }
