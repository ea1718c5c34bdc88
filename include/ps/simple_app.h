/**
 * \file simple_app.h
 * \brief SimpleApp: (int head, string body) request / response between nodes.
 *
 * API parity: reference include/ps/simple_app.h:14-184. Unlike the reference
 * (whose Request CHECK-fails for any receiver other than kServerGroup, SURVEY
 * appendix C) a request may address any node id or group id, and the tracker
 * expects one answer per addressed *instance*.
 */
#ifndef PS_SIMPLE_APP_H_
#define PS_SIMPLE_APP_H_
#include <functional>
#include <string>
#include <utility>
#include "ps/internal/message.h"
#include "ps/internal/postoffice.h"

namespace ps {

/*! \brief what a SimpleApp handler sees of a message */
struct SimpleData {
  /*! \brief application-defined opcode */
  int head = 0;
  /*! \brief application-defined payload */
  std::string body;
  /*! \brief node id of the peer */
  int sender = 0;
  /*! \brief request timestamp (echoed by the response) */
  int timestamp = 0;
  /*! \brief customer the message belongs to */
  int customer_id = 0;
};

class SimpleApp {
 public:
  using Handle = std::function<void(const SimpleData& recved, SimpleApp* app)>;

  /*!
   * \param app_id must match between the communicating apps
   * \param customer_id locally unique id of this endpoint
   * \param postoffice instance to attach to (default: the process's first one)
   */
  explicit SimpleApp(int app_id, int customer_id, Postoffice* postoffice = nullptr) : SimpleApp() {
    postoffice_ = postoffice != nullptr ? postoffice : Postoffice::Get();
    obj_ = new Customer(app_id, customer_id,
                        [this](const Message& m) { this->Process(m); }, postoffice_, false);
    obj_->Start();  // after the assignment: handlers reach the customer through obj_
  }
  virtual ~SimpleApp() {
    delete obj_;
    obj_ = nullptr;
  }

  /*! \brief send (head, body) to node / group `recv_id`; returns the timestamp to Wait() on */
  virtual int Request(int req_head, const std::string& req_body, int recv_id) {
    const auto& targets = postoffice_->GetNodeIDs(recv_id);
    const int ts = obj_->NewRequest(recv_id, static_cast<int>(targets.size()));
    Message out = Envelope(req_head, req_body, ts, /*request=*/true, obj_->customer_id());
    for (int node : targets) {
      out.meta.recver = node;
      postoffice_->van()->Send(out);
    }
    return ts;
  }

  /*! \brief block until every addressee of `timestamp` has responded */
  virtual void Wait(int timestamp) { obj_->WaitRequest(timestamp); }

  /*! \brief answer `recv_req`, echoing its head */
  virtual void Response(const SimpleData& recv_req, const std::string& res_body = "") {
    Message out = Envelope(recv_req.head, res_body, recv_req.timestamp, /*request=*/false,
                           recv_req.customer_id);
    out.meta.recver = recv_req.sender;
    postoffice_->van()->Send(out);
  }

  virtual void set_request_handle(const Handle& h) {
    CHECK(h) << "invalid request handle";
    on_request_ = h;
  }
  virtual void set_response_handle(const Handle& h) {
    CHECK(h) << "invalid response handle";
    on_response_ = h;
  }
  virtual Customer* get_customer() { return obj_; }
  /*!
   * \brief handle this app's messages on the van's receive thread instead of the customer
   *        thread whenever the queue is idle (extension; see Customer::set_inline_dispatch for
   *        the rules: handlers and callbacks must never wait for the network)
   */
  void set_inline_dispatch(bool on) { obj_->set_inline_dispatch(on); }

 protected:
  /*! \brief for subclasses (KVWorker / KVServer) that build their Customer themselves */
  SimpleApp() : obj_(nullptr), postoffice_(nullptr) {
    on_request_ = [](const SimpleData& req, SimpleApp* app) { app->Response(req); };
    on_response_ = [](const SimpleData&, SimpleApp*) {};
  }

  /*! \brief dispatch a received simple_app message to the installed handlers */
  virtual void Process(const Message& msg) {
    SimpleData d;
    d.head = msg.meta.head;
    d.body = msg.meta.body;
    d.sender = msg.meta.sender;
    d.timestamp = msg.meta.timestamp;
    d.customer_id = msg.meta.customer_id;
    const Handle& h = msg.meta.request ? on_request_ : on_response_;
    CHECK(h);
    h(d, this);
  }

  Customer* obj_;
  Postoffice* postoffice_;

 private:
  Message Envelope(int head, const std::string& body, int ts, bool request, int customer_id) {
    Message m;
    m.meta.simple_app = true;
    m.meta.request = request;
    m.meta.head = head;
    m.meta.body = body;
    m.meta.timestamp = ts;
    m.meta.app_id = obj_->app_id();
    m.meta.customer_id = customer_id;
    return m;
  }

  Handle on_request_;
  Handle on_response_;
};

}  // namespace ps
#endif  // PS_SIMPLE_APP_H_
