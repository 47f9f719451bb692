/* TEST INFRASTRUCTURE ONLY -- empty stand-in for <boost/thread.hpp>
 * (included by psk_soft_base.h:24; nothing from it is used on the demod path). */
#ifndef ORACLE_STUB_BOOST_THREAD_HPP
#define ORACLE_STUB_BOOST_THREAD_HPP
#endif
